// TEST INFRASTRUCTURE: see cuda_fp16.h.
#pragma once
#include <cstdint>
struct __nv_bfloat16 { uint16_t x; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
