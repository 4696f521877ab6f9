#!/usr/bin/env python
"""Summary of the REFERENCE's sample animator (samples/ozz/assets/animator.json; only readable where /root/reference exists):
state names, their clips with blend positions, blend type, threshold, tween duration, and the transitions.
Writes tests/golden/sample_animator.json; tests/test_golden.py requires the definition tests/cpp/animator_test.cpp builds
(and tests/test_host_shim.py's clip list) to be this one."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    src = json.load(open("/root/reference/samples/ozz/assets/animator.json"))
    states = {}
    for name, st in src["states"].items():
        if "elements" in st:
            clips = [[e["animation"], e["value"]] for e in st["elements"]]
        else:
            clips = [[st["animation"], [0, 0]]]
        states[name] = {"clips": clips, "blend": st.get("blend", "Cartesian"), "threshold": st.get("threshold"),
                        "tween_duration": st.get("tween", {}).get("duration")}
    transitions = {k: v.get("duration") for k, v in src["transitions"].items()}
    out = {"source": "samples/ozz/assets/animator.json", "states": states, "transitions": transitions}
    path = os.path.join(ROOT, "tests", "golden", "sample_animator.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(path)


if __name__ == "__main__":
    main()
