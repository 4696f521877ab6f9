// oracle/ref_shims/darmok/protobuf/easing.pb.h -- TEST INFRASTRUCTURE (see oracle/dmk_oracle.h).
//
// Stand-in for the header protoc would generate from assets/protobuf/easing.proto, so that the reference's own
// include/darmok/easing.hpp can be compiled where it lies (oracle/Makefile, target `ref`).  Only what that header uses:
// the message class name and its nested enum with the numbering of the .proto file.
#pragma once

namespace darmok::protobuf {
class Easing {
public:
    enum Type : int {
        Linear = 0, Stepped = 1,
        QuadraticIn = 2, QuadraticOut = 3, QuadraticInOut = 4,
        CubicIn = 5, CubicOut = 6, CubicInOut = 7,
        QuarticIn = 8, QuarticOut = 9, QuarticInOut = 10,
        QuinticIn = 11, QuinticOut = 12, QuinticInOut = 13,
        SinusoidalIn = 14, SinusoidalOut = 15, SinusoidalInOut = 16,
        ExponentialIn = 17, ExponentialOut = 18, ExponentialInOut = 19,
        CircularIn = 20, CircularOut = 21, CircularInOut = 22,
        BounceIn = 23, BounceOut = 24, BounceInOut = 25,
        ElasticIn = 26, ElasticOut = 27, ElasticInOut = 28,
        BackIn = 29, BackOut = 30, BackInOut = 31
    };
};
}  // namespace darmok::protobuf
