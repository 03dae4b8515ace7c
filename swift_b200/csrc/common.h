// Host-side shared declarations for the swift_b200 C-ABI library (no torch types anywhere).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace swb {

// last error message of the calling thread (returned by swb200_last_error())
void set_error(const char* fmt, ...);
const char* get_error();

#define SWB_CHECK_CUDA(expr)                                                                         \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      swb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));          \
      return static_cast<int>(_e) ? static_cast<int>(_e) : -1;                                       \
    }                                                                                                \
  } while (0)

#define SWB_REQUIRE(cond, ...)                                                                       \
  do {                                                                                               \
    if (!(cond)) {                                                                                   \
      swb::set_error(__VA_ARGS__);                                                                   \
      return SWB_ERR_INVALID;                                                                        \
    }                                                                                                \
  } while (0)

constexpr int SWB_OK = 0;
constexpr int SWB_ERR_INVALID = -2;     // bad argument / unsupported configuration
constexpr int SWB_ERR_DRIVER = -3;      // driver entry point or tensor-map encode failure
constexpr int SWB_ERR_RESIDENCY = -4;   // the fused LayerNorm GEMM needs every CTA of its grid resident and the device cannot hold them

int num_sms();            // of the calling thread's current device

// One-time per-DEVICE initialisation (cudaFuncSetAttribute, occupancy queries) keyed by the current device id: a process
// that drives several GPUs must repeat it on each of them.  Usage:  static PerDevice<bool> done;  if (!done.get()) {...; done.set(true);}
constexpr int kMaxDevices = 64;
int current_device();
template <typename T>
struct PerDevice {
  T v[kMaxDevices] = {};
  T get() const { return v[current_device()]; }
  void set(T x) { v[current_device()] = x; }
};

// 16-bit (fp16 / bf16) row-major [rows, cols] with row pitch `pitch_elems` -> 2-D TMA descriptor with a
// [box_rows x box_cols] SWIZZLE_128B box (box_cols * 2 bytes must be 128)
int make_tmap_16bit_2d(CUtensorMap* out, const void* ptr, bool is_f16, uint64_t rows, uint64_t cols,
                       uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols);

struct GemmParams;
// D = A[M,K] * W[N,K]^T with a fused epilogue (gemm_sm100.cuh).
// tile: 1 = single CTA 128x176, 2 = CTA pair 256x176, 3 = CTA pair 256x352 (the SwiGLU weight packing depends on it);
// act_f16: the 16-bit operand format (A, W and 16-bit outputs) is fp16 instead of bf16.
int launch_gemm(int epi, int tile, int act_f16, const void* A, int lda, const void* W, int ldw, const GemmParams& p,
                cudaStream_t stream);

}  // namespace swb
