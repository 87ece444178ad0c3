// Host prelude for running the DEVICE FUNCTIONS of csrc/fm_form.cuh on the CPU under ASan / UBSan.  Tests only: a
// sanitiser + parity pass over the kernel SOURCE (tests/test_kernel_source_host.py assembles the translation unit from
// the real csrc files); it is not a CPU path of the product and nothing in fair_marl_b200 links it.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "fairmarl.h"

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __noinline__
struct Dim3 { int x = 0, y = 0, z = 0; };
static Dim3 blockIdx, blockDim, threadIdx;

static inline double __dadd_rn(double a, double b) { return a + b; }     // built with -ffp-contract=off
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int max(int a, int b) { return a > b ? a : b; }      // CUDA's global int overloads
static inline int min(int a, int b) { return a < b ? a : b; }
using std::fmaf;   // fabsf / fmaxf come from <cmath> in the global namespace
using std::exp; using std::fabs; using std::fmax; using std::fmin; using std::log1p; using std::sqrt; using std::tanh;
#define __restrict__
#define FM_DIV64(a, b) ((a) / (b))
#define FM_SQRT64 sqrt                                  // the device uses dsqrt_fast (same bits: correctly rounded)
namespace fm {                                          // hardware approximations of the device build (fm_device.cuh)
static inline float rsqrt_approx(float x) { return 1.0f / std::sqrt(x); }
static inline float ex2_approx(float x) { return std::exp2(x); }
}

namespace fm {
constexpr int INFO_F = 14;          // fm_device.cuh
constexpr int MAX_DRAWS = 4096;     // fm_device.cuh
}
