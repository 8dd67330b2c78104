// Force-included (nvcc -include) ahead of the reference translation unit when building
// oracle/_ref/libnbco_ref.so.  TEST INFRASTRUCTURE ONLY.
//
// The reference does not build as shipped with g++ 13 / nvcc 12.9:
//   * constants.cuh:110,120 use uint32_t without <cstdint>; reductions.cuh:70 and
//     main3.cu:742 use FLT_MAX without <cfloat>;
//   * helper_math.h:1566,1583 define a global lerp() that collides with C++20 std::lerp.
// Neither fix touches arithmetic.  <cmath> is pulled in *before* the rename so that
// std::lerp keeps its name and only the reference's helper gets a new one.
#pragma once
#include <cmath>
#include <cstdint>
#include <cfloat>
#include <cstdlib>
#include <numeric>
#include <algorithm>
#include <string>
#include <complex>
#include <random>
#include <atomic>
#include <thread>
#include <vector>
#include <iostream>
#include <fstream>
#include <chrono>
#include <bit>
#define lerp nbco_ref_helper_lerp
