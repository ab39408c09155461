// TEST INFRASTRUCTURE: storage-only stand-ins so that common.cuh parses on the host; no emulated kernel uses 16-bit data.
#pragma once
#include <cstdint>
struct __half { uint16_t x; };
struct __half2 { __half x, y; };
inline __half __float2half_rn(float) { return __half{0}; }
inline float __half2float(__half) { return 0.f; }
inline __half2 __floats2half2_rn(float, float) { return __half2{}; }
inline float2 __half22float2(__half2) { return float2{0.f, 0.f}; }
