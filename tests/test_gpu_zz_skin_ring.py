"""The skinning kernel's character loop with every block walking many characters (staging and replica rings wrap), every
output shape and vertex order, against the oracle: tests/skin_ring_check.py in a child process.  The default shape's loop is what
the full-size tests and bench.py have run on hardware since round 1; the other shapes, the producer-warp kernel and the
bone-clustered vertex order were verified on a B200 in round 2 (the same script runs on the CPU lockstep simulator,
tests/test_devsim.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_character_loop_of_every_skinning_shape_matches_the_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "skin_ring_check.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "cases match the oracle" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_four_vertices_per_thread_shape_matches_the_oracle():
    """DMK_SKIN_VPT=4: two lanes share an 8-vertex block (the second passes its first 16 bytes to the first by shuffles so that
    every store is a whole sector).  A diagnostic shape (measured slower, profiles/r02_skin_vpt4_experiment.jsonl); same checks."""
    env = dict(os.environ, DMK_SKIN_VPT="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "skin_ring_check.py"), "--quick"], capture_output=True, text=True, timeout=300, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "cases match the oracle" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_bulk_store_output_path_matches_the_oracle():
    """DMK_SKIN_BULK_STORE=1: the clustered mesh stored in the strided layout (pitch a multiple of 64), results staged in shared memory
    and sent by cp.async.bulk shared -> global per round of 32 vertices.  A diagnostic shape (measured slower: the SM's bulk-copy
    engine is already half busy with the palette replicas, profiles/r02_skin_bulk_store_experiment.txt); same checks."""
    env = dict(os.environ, DMK_SKIN_BULK_STORE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "skin_ring_check.py"), "--quick"], capture_output=True, text=True, timeout=300, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "cases match the oracle" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_static_split_of_the_characters_matches_the_oracle():
    """DMK_SKIN_STATIC_SPLIT=1: every block takes a fixed share of the characters (the schedule before the per-tile counters; kept as
    the A/B side of profiles/r02_skin_dynamic_assignment.txt); same checks."""
    env = dict(os.environ, DMK_SKIN_STATIC_SPLIT="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "skin_ring_check.py"), "--quick"], capture_output=True, text=True, timeout=300, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "cases match the oracle" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
