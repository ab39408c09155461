// TEST INFRASTRUCTURE: host stand-in for the few fp16 conversions common.cuh uses (g++'s _Float16: IEEE binary16,
// round-to-nearest-even conversions, as __float2half_rn).
#pragma once
#include <cstdint>
#include <cstring>
struct __half { uint16_t x; };
struct __half2 { __half x, y; };
inline __half __float2half_rn(float f) { _Float16 h = (_Float16)f; __half r; std::memcpy(&r.x, &h, 2); return r; }
inline float __half2float(__half v) { _Float16 h; std::memcpy(&h, &v.x, 2); return (float)h; }
inline __half2 __floats2half2_rn(float a, float b) { return __half2{__float2half_rn(a), __float2half_rn(b)}; }
inline float2 __half22float2(__half2 v) { return float2{__half2float(v.x), __half2float(v.y)}; }
