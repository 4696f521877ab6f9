// oracle/ref_easing.cpp -- TEST INFRASTRUCTURE (see oracle/dmk_oracle.h).
//
// C entry point around the REFERENCE's own easing curves: this file includes /root/reference/include/darmok/easing.hpp
// where it lies (nothing of it is copied) and is compiled into oracle/_ref/libref_easing.so by `make -C oracle ref`.
// Easing::apply is what the transition blend (src/skeleton_ozz.cpp:712) and the tweened blend position (:566, :574)
// evaluate; the library pins the oracle's orc::ease and the shim's ease01 (tests/golden/make_easing_golden.py).
#include <darmok/easing.hpp>

extern "C" __attribute__((visibility("default"))) float ref_easing_apply(int type, float position, float start, float end) {
    return darmok::Easing::apply<float>(static_cast<darmok::EasingType>(type), position, start, end);
}
