// tcgen05 implicit-GEMM convolution engine for sm_100a (B200).
//
// Computes the 3x3 / 1x1 convolutions and linears of the CodeFormer hot path
//   nn.Conv2d call sites      /root/reference/basicsr/archs/vqgan_arch.py:120,132,147-151,173-200,243,266,292,314
//   Fuse_sft convs, Linear    /root/reference/basicsr/archs/codeformer_arch.py:104-106,141-149,183,192
// as GEMMs  D[M = 128 output pixels, N = Cout tile] += A[M, K] * B[N, K]^T  with K = taps * Cin, on the
// 5th-generation tensor cores:
//   * operands are error-compensated fp16 pairs  x = hi + lo  (hi = fp16(x), lo = fp16(x - hi)); the products
//     hi*hi + hi*lo + lo*hi are formed with TWO tcgen05.mma kind::f16 per k-step: A_hi x [B_hi;B_lo] (one N = 2*BN
//     operand, the hi and lo weight tiles are adjacent in smem) and A_lo x B_hi, into a "main" and a "cross" half of
//     an fp32 TMEM slot (>= 21 effective mantissa bits; the 1e-3 parity bar needs >= 16);
//   * tcgen05.mma truncates when it adds into its accumulator, so a TMEM slot only receives `chunk` k-blocks before the
//     epilogue warps fold it into fp32 registers with round-to-nearest adds (ring of 2-4 slots: cfull/cempty);
//   * A tiles are fetched by TMA straight from the NHWC operand planes: one 4-D box {64 ch, BW, BH, 1} per filter tap at
//     shifted (x+s-1, y+r-1) coordinates -- out-of-bounds rows/cols are zero-filled by the TMA unit, which *is* the conv
//     padding; the box lands in shared memory as 128 rows x 128 B in the 128B-swizzled K-major layout the UMMA
//     descriptor expects (no im2col buffer anywhere).  Stride-2 Downsample = TMA traversal strides (elementStrides 2);
//     Upsample = four 2x2 parity convs on the low-res planes with pre-summed weights;
//   * B tiles ([tap][Cout][Cin] fp16) by 3-D TMA boxes {64, BN, 1};
//   * warp-specialised persistent CTAs (1 per SM): warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), both kept
//     converged with elect.sync-predicated issue; warps 2..9 = epilogue (tcgen05.ld -> fold -> per-warp swizzled smem
//     transpose -> bias / residual / activation / SFT -> coalesced fp32 NHWC stores, optional fp16 hi/lo planes for a
//     raw-input consumer, GroupNorm(32) partial sums);
//   * default engine = CTA PAIR + HALO (PAIR / HALO template flags): two CTAs of a cluster (one TPC) own adjacent 8x16-pixel
//     tiles and share the weights through tcgen05.mma.cta_group::2 (M = 256): no single-CTA instruction floor, each CTA
//     stages half of every B operand, commits are multicast to both CTAs, both CTAs' TMA bytes are counted on the leader's
//     mbarrier; the (10x18) input patch of a 64-channel block is fetched ONCE and the taps read it through row-shifted
//     128B-swizzled descriptors (see the comments at TcCfg and tc_geometry; probes in tools/umma_*.py);
//   * in-kernel operand transform (XF, the default for every GroupNorm(+SiLU) consumer): the patch of a 64-channel block
//     arrives as the fp32 ACTIVATION itself (own loader warp, two 32-channel TMA boxes) and 4 (BN = 128) or 8 (BN = 64)
//     transform warps apply GroupNorm-affine + SiLU + the fp16 hi/lo split in place before the MMAs read it -- no operand
//     planes in HBM, no separate preparation pass (tc_can_xform says where it is used); GEN = the same for any H x W with
//     reflection / replicate padding (ParseNet, RRDBNet); the kernel also runs the attention GEMMs (bmm_tc);
//   * every kernel of the forward is launched with programmatic stream serialization (PDL, kernels.cuh); few-tile launches of
//     wide layers use 64-wide n-tiles (conv_tc()); VectorQuantizer.forward is its own one-kernel pipeline (vq_fused_kernel).
// Diagnostics are BUILD options (-DCFB_TC_STAMPS=1): measured on B200, stamp tests inside the role loops and even two unused
// fields in the kernel parameter block slow every conv kernel by 10-25 % (profiles/round2_ab_builds.txt) -- keep TcParams compact.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "conv_tc.cuh"

namespace cfb {

// ------------------------------------------------------------------------------------------------------
// weight split: OIHW fp32 -> [tap][Cout][Cin] fp16 hi/lo of (w * 2^k), k chosen so max|w| lands in [2^13,2^14)
// (keeps `lo` out of the fp16 subnormal range); scale_slot[0] = |w|max bits, scale_slot[1] = 2^-k for the epilogue
// ------------------------------------------------------------------------------------------------------
__global__ void tc_absmax_kernel(const float* __restrict__ w, int64_t total, unsigned* __restrict__ slot) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));
}

__global__ void tc_split_weights_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo,
                                        int Cout, int Cin, int k, float* __restrict__ slot) {
  const float amax = __uint_as_float(reinterpret_cast<const unsigned*>(slot)[0]);
  int e = 0;
  if (amax > 0.f && isfinite(amax)) frexpf(amax, &e);
  const float scale = exp2f((float)(14 - e));
  const int64_t total = (int64_t)Cout * Cin * k * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int tap = (int)(i / ((int64_t)Cout * Cin));
    const float v = w[((int64_t)co * Cin + ci) * k * k + tap] * scale;
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) slot[1] = exp2f((float)(e - 14));
}

int tc_split_weights(const float* oihw, __half* hi, __half* lo, int Cout, int Cin, int k, float* scale_slot,
                     cudaStream_t st) {
  const int64_t total = (int64_t)Cout * Cin * k * k;
  const int64_t blocks = (total + 255) / 256;
  const unsigned g = (unsigned)(blocks > 1024 ? 1024 : blocks);
  CFB_CUDA(cudaMemsetAsync(scale_slot, 0, 2 * sizeof(float), st));
  tc_absmax_kernel<<<g, 256, 0, st>>>(oihw, total, reinterpret_cast<unsigned*>(scale_slot));
  CFB_LAUNCH_CHECK();
  tc_split_weights_kernel<<<g, 256, 0, st>>>(oihw, hi, lo, Cout, Cin, k, scale_slot);
  CFB_LAUNCH_CHECK();
  return 0;
}

// Upsample (nearest x2) followed by a 3x3 conv == four 2x2 convs on the LOW-resolution tensor, one per output parity
// (py,px): out[2y+py][2x+px] = sum_{dy,dx in {0,1}} W'[py][px][dy][dx] . in[y+dy+py-1][x+dx+px-1], where W' pre-sums the 3x3
// taps that read the same source pixel (rows: py=0 -> {r0 | r1+r2}, py=1 -> {r0+r1 | r2}; same for columns).  2.25x fewer
// MACs and the operand planes stay at the low resolution.  Sums are formed in fp32 before the hi/lo split.
__device__ __forceinline__ void up4_range(int parity, int d, int& lo, int& hi) {
  if (parity == 0) { lo = d == 0 ? 0 : 1; hi = d == 0 ? 0 : 2; }
  else { lo = d == 0 ? 0 : 2; hi = d == 0 ? 1 : 2; }
}
__device__ __forceinline__ float up4_weight(const float* __restrict__ w, int co, int ci, int Cin, int tap16) {
  const int ph = tap16 >> 2, py = ph >> 1, px = ph & 1, dy = (tap16 >> 1) & 1, dx = tap16 & 1;
  int r0, r1, s0, s1;
  up4_range(py, dy, r0, r1);
  up4_range(px, dx, s0, s1);
  float v = 0.f;
  for (int r = r0; r <= r1; ++r)
    for (int q = s0; q <= s1; ++q) v += w[(((int64_t)co * Cin + ci) * 3 + r) * 3 + q];
  return v;
}
__global__ void tc_absmax_up4_kernel(const float* __restrict__ w, int Cout, int Cin, unsigned* __restrict__ slot) {
  const int64_t total = (int64_t)16 * Cout * Cin;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin), co = (int)((i / Cin) % Cout), tap = (int)(i / ((int64_t)Cout * Cin));
    m = fmaxf(m, fabsf(up4_weight(w, co, ci, Cin, tap)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));
}
__global__ void tc_split_up4_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, int Cout,
                                    int Cin, float* __restrict__ slot) {
  const float amax = __uint_as_float(reinterpret_cast<const unsigned*>(slot)[0]);
  int e = 0;
  if (amax > 0.f && isfinite(amax)) frexpf(amax, &e);
  const float scale = exp2f((float)(14 - e));
  const int64_t total = (int64_t)16 * Cout * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin), co = (int)((i / Cin) % Cout), tap = (int)(i / ((int64_t)Cout * Cin));
    const float v = up4_weight(w, co, ci, Cin, tap) * scale;
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) slot[1] = exp2f((float)(e - 14));
}
int tc_split_weights_up4(const float* oihw3x3, __half* hi, __half* lo, int Cout, int Cin, float* scale_slot, cudaStream_t st) {
  const int64_t total = (int64_t)16 * Cout * Cin;
  const int64_t blocks = (total + 255) / 256;
  const unsigned g = (unsigned)(blocks > 1024 ? 1024 : blocks);
  CFB_CUDA(cudaMemsetAsync(scale_slot, 0, 2 * sizeof(float), st));
  tc_absmax_up4_kernel<<<g, 256, 0, st>>>(oihw3x3, Cout, Cin, reinterpret_cast<unsigned*>(scale_slot));
  CFB_LAUNCH_CHECK();
  tc_split_up4_kernel<<<g, 256, 0, st>>>(oihw3x3, hi, lo, Cout, Cin, scale_slot);
  CFB_LAUNCH_CHECK();
  return 0;
}

__device__ void report_overflow();       // status word, defined with the barrier helpers below

// ------------------------------------------------------------------------------------------------------
// operand preparation: fp32 NHWC (+ fused GroupNorm affine, SiLU, nearest x2) -> fp16 hi / lo NHWC planes
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tc_silu(float x) { return x / (1.f + expf(-x)); }

// One block = PB consecutive output pixels of ONE image; a thread keeps the same 8-channel slice for all its pixels, so
// the per-(n,c) GroupNorm scale/shift is loaded once per block instead of once per element.
__global__ void __launch_bounds__(256) tc_prep_kernel(const float* __restrict__ in, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, int act, int up, int N, int H, int W,
                                                      int C, int PB, __half* __restrict__ hi, __half* __restrict__ lo) {
  const int C8 = C >> 3;
  const int Hp = H << up, Wp = W << up;
  const int64_t img_px = (int64_t)Hp * Wp;
  const int64_t pix0 = (int64_t)blockIdx.x * PB;          // PB divides Hp*Wp: the block stays inside image n
  const int n = (int)(pix0 / img_px);
  const int c = (threadIdx.x % C8) * 8;
  const int pstep = 256 / C8;
  float sv[8], hv[8];
  if (scale) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + (int64_t)n * C + c));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + (int64_t)n * C + c + 4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + (int64_t)n * C + c));
    const float4 h1 = __ldg(reinterpret_cast<const float4*>(shift + (int64_t)n * C + c + 4));
    sv[0] = s0.x; sv[1] = s0.y; sv[2] = s0.z; sv[3] = s0.w; sv[4] = s1.x; sv[5] = s1.y; sv[6] = s1.z; sv[7] = s1.w;
    hv[0] = h0.x; hv[1] = h0.y; hv[2] = h0.z; hv[3] = h0.w; hv[4] = h1.x; hv[5] = h1.y; hv[6] = h1.z; hv[7] = h1.w;
  }
  float vmax = 0.f;
  for (int pl = threadIdx.x / C8; pl < PB; pl += pstep) {
    const int64_t pix = pix0 + pl;
    const int64_t rem = pix - (int64_t)n * img_px;
    const int oy = (int)(rem / Wp), ox = (int)(rem - (int64_t)oy * Wp);
    const float* src = in + (((int64_t)n * H + (oy >> up)) * W + (ox >> up)) * C + c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sv[j], hv[j]);
    }
    if (act == IN_SILU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tc_silu(v[j]);
    }
    __align__(16) __half hh[8];
    __align__(16) __half ll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      vmax = fmaxf(vmax, fabsf(v[j]));
      hh[j] = __float2half_rn(v[j]);
      ll[j] = __float2half_rn(v[j] - __half2float(hh[j]));
    }
    *reinterpret_cast<uint4*>(hi + pix * C + c) = *reinterpret_cast<const uint4*>(hh);
    *reinterpret_cast<uint4*>(lo + pix * C + c) = *reinterpret_cast<const uint4*>(ll);
  }
  if (vmax > 65504.f) report_overflow();       // fp16 operand range guard: reported through the status word, never silent
}

// torch.cat([enc_feat, dec], dim=1) of Fuse_sft_block (codeformer_arch.py:152) written directly as RAW fp16 hi/lo operand
// planes: the fused ResBlock reads the concatenation only through the tensor engine (conv1 transforms it in-kernel, the
// 1x1 conv_out takes it raw) and its GroupNorm statistics come from the sources' partial sums, so no fp32 copy is needed.
__global__ void __launch_bounds__(256) concat_planes_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            __half* __restrict__ hi, __half* __restrict__ lo, int64_t pixels,
                                                            int Ca, int Cb) {
  pdl_launch_dependents();
  pdl_wait();
  const int C8 = (Ca + Cb) >> 3;
  const int64_t total = pixels * C8;
  float vmax = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t px = i / C8;
    const int c = (int)(i - px * C8) * 8;
    const float* src = c < Ca ? a + px * Ca + c : b + px * Cb + (c - Ca);
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    __align__(16) __half hh[8];
    __align__(16) __half ll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      vmax = fmaxf(vmax, fabsf(v[j]));
      hh[j] = __float2half_rn(v[j]);
      ll[j] = __float2half_rn(v[j] - __half2float(hh[j]));
    }
    *reinterpret_cast<uint4*>(hi + px * (Ca + Cb) + c) = *reinterpret_cast<const uint4*>(hh);
    *reinterpret_cast<uint4*>(lo + px * (Ca + Cb) + c) = *reinterpret_cast<const uint4*>(ll);
  }
  if (vmax > 65504.f) report_overflow();
}
int concat_planes(const float* a, const float* b, void* planes, int64_t pixels, int Ca, int Cb, cudaStream_t st) {
  CFB_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0, "concat_planes: channel counts must be multiples of 8");
  if (pixels == 0) return 0;
  const size_t plane = ((size_t)pixels * (Ca + Cb) * 2 + 1023) / 1024 * 1024;
  const int64_t total = pixels * ((Ca + Cb) >> 3);
  const int64_t blocks = (total + 255) / 256;
  CFB_LAUNCH_PDL(concat_planes_kernel, dim3((unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks)), dim3(256), 0, st, a, b, (__half*)planes,
                 (__half*)((char*)planes + plane), pixels, Ca, Cb);
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait without a trap.  A protocol bug (or an injected fault) must surface as a Python exception, never as a hung
// GPU and never as a sticky context error (SURVEY.md section 8(b) "Errors": the reference's callers catch RuntimeError and
// fall back to the input face, inference_codeformer.py:209-211).  On time-out the waiting thread raises the device-wide
// abort flag and reports through the host-mapped status word; every other wait loop sees the flag after its next failed
// try_wait; a warp that has seen it leaves its role loop at once (no further TMA / MMA / barrier traffic), all warps meet at
// the kernel's tear-down, the kernel exits normally and the host turns the status word into an error (runtime.cu:
// async_status_check).  The results of that launch are garbage by contract.
__device__ unsigned g_abort = 0;                 // per device: set on a barrier time-out, cleared by the host when it reports it
__device__ unsigned* g_status_host = nullptr;    // host-mapped status word (bit 0: barrier time-out, bit 1: fp16 operand overflow)
__device__ long long g_wait_limit = 4000000000LL;   // cycles (~2 s); the fault-injection test lowers it

__device__ __noinline__ void mbar_timeout() {
  atomicExch(&g_abort, 1u);
  if (g_status_host) { atomicOr_system(g_status_host, CFB_STATUS_TIMEOUT); __threadfence_system(); }
}
__device__ __noinline__ void report_overflow() {
  if (g_status_host) { atomicOr_system(g_status_host, CFB_STATUS_OVERFLOW); __threadfence_system(); }
}
// RELAX_NS > 0: a producer / helper role whose wake-up latency is not critical sleeps between polls, leaving the issue slots
// to the working warps of its scheduler (a failed try_wait comes back after a few hundred cycles, whatever the hint says:
// in the round-2 profile the poll loops of 19 warps were 40 % of all executed instructions).  The abort flag and the
// time-out clock are only looked at every 128 failed polls.
template <int RELAX_NS = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, bool& aborted) {
  uint32_t done = 0;
  uint32_t polls = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (done) break;
    if (RELAX_NS > 0) __nanosleep(RELAX_NS);
    if ((++polls & 127u) == 0u) {
      if (*(volatile unsigned*)&g_abort) { aborted = true; break; }
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > *(volatile long long*)&g_wait_limit) { mbar_timeout(); aborted = true; break; }
    }
  }
  aborted = __any_sync(0xffffffffu, aborted);      // the whole (converged) warp takes the same decision
}
// diagnostics kernels: plain bounded wait
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (!done && clock64() - t0 > *(volatile long long*)&g_wait_limit) { mbar_timeout(); break; }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// One leader lane of a fully converged warp (same lane every time).  Keeping the role warps converged and predicating
// only the issue instructions lets ptxas hold descriptors / addresses in uniform registers; an `if (lane == 0)` region
// instead forces a per-instruction ELECT/branch waterfall around every UTCHMMA (measured: issue-bound at N=64).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// descriptors as {lo32 = address field | LBO, hi32 = SBO | version | layout}: only lo changes between MMAs
__device__ __forceinline__ void tc_mma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
constexpr uint32_t DESC_HI_SW128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B

// K-major, 128B-swizzled operand tile (rows of 128 B, 8-row atoms 1024 B apart): cute::UMMA::SmemDescriptor
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}


// ---- CTA-pair (cta_group::2) plumbing ----
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {      // shared::cta -> shared::cluster of `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Remote arrival after this warp's shared-memory writes (+ fence.proxy.async): CTA-scope release, as the data it publishes is
// this CTA's own shared memory, read by this SM's tensor core -- the remote thread only issues the instruction.  The
// .release.cluster form costs MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR, ~1000 cycles per arrival (profiles/round2_c64_xf_full).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// Arrival that orders nothing but itself: for barriers that only hand TMEM back to the MMA issuer.  The tcgen05.ld results are
// already in registers (tcgen05.wait::ld) and tcgen05.fence::before_thread_sync orders the tensor-memory accesses; the default
// .release form additionally drains every global store this thread still has in flight (MEMBAR + ERRBAR: 30 % of the epilogue
// warps' time in profiles/round2_c64_xf_full, and it delays the accumulator slot the MMA warp is waiting for).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// wait on a LOCAL barrier whose arrivals may come from the peer CTA (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cl(uint32_t bar, uint32_t parity, bool& aborted) {
  uint32_t done = 0;
  uint32_t polls = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (done) break;
    if ((++polls & 127u) == 0u) {
      if (*(volatile unsigned*)&g_abort) { aborted = true; break; }
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > *(volatile long long*)&g_wait_limit) { mbar_timeout(); aborted = true; break; }
    }
  }
  aborted = __any_sync(0xffffffffu, aborted);
}
// TMA loads of a CTA pair: data lands in the issuing CTA's shared memory, the byte count is signalled on `cluster_bar`,
// which may live in the peer (leader) CTA
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ------------------------------------------------------------------------------------------------------
// fused operand transform: per-patch worker of the transform warps (see conv_tc_kernel, XF)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_ftz(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {      // {hi:16 | lo:16}, round to nearest even
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// float(h) - c as ONE mixed-precision instruction (FHADD); h = fp16 bits
__device__ __forceinline__ float f16_minus_f32(uint32_t h, float c) {
  float r;
  asm("{\n\t.reg .b16 hh;\n\tcvt.u16.u32 hh, %1;\n\tsub.rn.f32.f16 %0, hh, %2;\n\t}" : "=f"(r) : "r"(h), "f"(c));
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// float4 forms for the epilogue's transpose patches.  The dynamic shared-memory base is realigned through an integer cast, so
// plain pointer accesses compile to GENERIC ld/st (LD.E / ST.E with 64-bit address arithmetic); these stay in the shared window.
// No "memory" clobber: volatile asms keep their mutual order and every write -> read hand-off of a patch crosses a
// __syncwarp(), while the compiler stays free to move the residual / SFT global loads of a row batch ahead of them.
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}
// 8 raw values -> y = act(x * sc + sh) -> fp16 hi / lo words.  MODE 2: affine + SiLU, 1: affine, 0: plain split.
// SiLU = y / (1 + 2^(-y log2 e)) with ex2.approx / rcp.approx (~2^-21 relative: below the hi/lo operand error of 2^-22..2^-21).
// lo = rn(y - hi) is formed as -(hi - y) with one mixed-precision subtract per element and a sign flip of the packed pair.
template <int MODE>
__device__ __forceinline__ void xf_chunk(const uint4& a, const uint4& b, const float (&sc)[8], const float (&sh)[8], uint4& hv,
                                         uint4& lv, float& amax) {
  const uint32_t raw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float y[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float v = __uint_as_float(raw[2 * q + e]);
      if (MODE >= 1) v = fmaf(v, sc[2 * q + e], sh[2 * q + e]);
      if (MODE == 2) v = v * rcp_ftz(1.f + ex2_ftz(v * -1.4426950408889634f));
      y[e] = v;
    }
    amax = fmaxf(amax, fmaxf(fabsf(y[0]), fabsf(y[1])));
    hw[q] = pack_f16x2(y[0], y[1]);
    const float d0 = f16_minus_f32(hw[q] & 0xffffu, y[0]);      // hi - y = -lo
    const float d1 = f16_minus_f32(hw[q] >> 16, y[1]);
    lw[q] = pack_f16x2(d0, d1) ^ 0x80008000u;
  }
  hv = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  lv = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
// One patch: RPP rows per pass (8 lanes per row), two passes in flight per iteration.
template <int MODE, int RPP, int NPASS, int PW, int ROWS>
__device__ __forceinline__ void xf_patch(uint32_t src_base, uint32_t hi_base, uint32_t lo_base, int c0, int j, int rsub,
                                         const float (&sc)[8], const float (&sh)[8], bool border, int y0, int x0, int Hin,
                                         int Win, float& amax) {
#pragma unroll
  for (int it = 0; it < NPASS; it += 2) {
    uint4 a[2], b[2];
    bool act[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = rsub + (it + u) * RPP;
      act[u] = (it + u < NPASS) && (r < ROWS);                 // warp-uniform: a warp owns 4 consecutive rows and ROWS % 4 == 0
      if (act[u]) {
        const uint32_t row = src_base + (uint32_t)r * 128u;
        const uint32_t sw = (row >> 7) & 7u;                    // swizzle phase = absolute address bits [7,10)
        a[u] = lds128(row + (((uint32_t)c0 ^ sw) << 4));
        b[u] = lds128(row + (((uint32_t)(c0 + 1) ^ sw) << 4));
      }
    }
    __syncwarp();                                              // every lane of the row has loaded before any lane stores
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (act[u]) {
        const int r = rsub + (it + u) * RPP;
        uint4 hv, lv;
        xf_chunk<MODE>(a[u], b[u], sc, sh, hv, lv, amax);
        if (border) {
          const int py = r / PW, px = r - py * PW;
          if (!((unsigned)(y0 + py) < (unsigned)Hin && (unsigned)(x0 + px) < (unsigned)Win)) {
            hv = make_uint4(0u, 0u, 0u, 0u);
            lv = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        const uint32_t hrow = hi_base + (uint32_t)r * 128u, lrow = lo_base + (uint32_t)r * 128u;
        sts128(hrow + ((((uint32_t)j) ^ ((hrow >> 7) & 7u)) << 4), hv);
        sts128(lrow + ((((uint32_t)j) ^ ((lrow >> 7) & 7u)) << 4), lv);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------
// Phase stamps (tools/tc_stamps.py) are a BUILD option: -DCFB_TC_STAMPS=1.  Measured on B200 (profiles/round2_ab_builds.txt):
// the stamp tests inside the role loops cost 6-13 % of every conv kernel even with a null stamp buffer, and the mere presence
// of the stamp pointer (+ one more int) in this struct another 3-15 %, so production builds omit both.
#ifndef CFB_TC_STAMPS
#define CFB_TC_STAMPS 0
#endif
struct TcParams {
  int N, Ho, Wo, Cout;
  int taps, pad, stride;  // 9/1/1 (3x3 'same'), 1/0/1 (1x1), 9/0/2 (Downsample: pad right/bottom = TMA OOB zero fill)
  int BW, BH;             // pixel tile = BH rows x BW cols = 128
  int tiles_x, tiles_y;   // per image
  int m_tiles, n_tiles;
  int kblocks;            // Cin / 64
  int chunk;              // k-blocks accumulated in TMEM before the partial sum is folded into registers
  int a_c0, b_c0;         // channel offsets of the A / B operand inside their planes (batched-GEMM mode)
  int b_batched;          // B operand is a per-image activation plane: third TMA coordinate = image index, not the tap
  // batched GEMM over (image, head) pairs (multi-head attention): the "image" index nb of an m-tile is n * heads + h
  int heads;              // 1: plain
  int a_c_head, b_c_head; // channel offset per head inside the A / B planes
  int a_img_per_head;     // A planes hold one image per (n, h) (softmax probabilities) instead of one per n
  int b_r_head;           // row offset per head inside the B planes (V^T: rows = h*64 + c)
  int out_per_head;       // 1: out is [n*heads + h][256][Cout]; 0: out is [n][256][Cout] and head h owns columns [h*o_c_head, ..)
  int o_c_head;
  int up4;                // Upsample as four 2x2 convs: m-tile = (low-res tile, output parity), 4 taps, weights [16][Cout][Cin]
  int PW, PH;             // halo engine: input patch (BW+k-1) x (BH+k-1) pixels fetched once per 64-channel block
  // GEN variant (XF only; ParseNet / RRDBNet): true image sizes with ragged tiles, padding mode of the halo patch, output
  // placement into a wider buffer, a second (scaled) residual, stride-2 by subsampling
  int Hin, Win;           // true input height / width (tiles are ceil-divided; out-of-image patch pixels follow pad_mode)
  int pad_mode;           // 0 zero, 1 reflect (ReflectionPad2d), 2 replicate (reflection padding of a nearest-x2 upsampled tensor)
  int sub;                // 1: keep only the even output positions (3x3 stride-2 pad-1 conv == its stride-1 result subsampled)
  int out_pitch, out_c0, cout_valid;   // destination: channels per pixel, channel offset, number of real output channels
  int res_pitch;          // channels per pixel of `residual`
  const float* residual2; // out = (conv + bias + residual) * post_scale + residual2
  int res2_pitch;
  float post_scale;
  const float* vq_e2; const float* vq_z2; float2* vq_cand; double* vq_dpart;   // VectorQuantizer argmin epilogue (see ConvArgs)
  int fault;              // test hook (cfb_debug_inject_fault): CTA 0 drops the weight load of its first stage -> barrier time-out
  int xform;              // XF kernel variant requested (in_scale may be null: raw split)
  int a_split;            // XF: k-blocks [0, a_split) are read from fp32 source 0 (tmA_hi), the rest from source 1 (tmA_lo)
  const float* in_scale;    // XF: per-(n, cin) affine of the fused operand transform (GroupNorm folded), and its activation
  const float* in_shift;
  int in_act;
  const float* bias;
  const float* residual;
  int out_act;
  const float* sft_dec;
  const float* sft_scale;
  float sft_w;
  const float* wscale_inv;  // device scalar: 2^-k of the weight split
  float* out;
  __half* pl_hi;            // optional: `out` again as fp16 hi / lo planes (operand of a following raw-input conv)
  __half* pl_lo;
  float* gn_part;           // optional GroupNorm(32) partial sums of `out`: [m_tile*4 + warp][32][2]
  int gn_cpg;               // channels per group = Cout/32
#if CFB_TC_STAMPS
  long long* dbg;           // diagnostics (cfb_debug_set_stamps): CTA 0 writes clock64() at its role hand-offs; null in production
#endif
};
// phase stamps of CTA 0 (tools/tc_stamps.py): one lane per role writes the SM cycle counter
#if CFB_TC_STAMPS
#define TC_STAMP(i) do { if (p.dbg != nullptr && blockIdx.x == 0) p.dbg[i] = clock64(); } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#endif

constexpr int TC_EPI_WARPS = 8;                       // 4 TMEM lane quadrants x 2 column halves
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // warp0 TMA, warp1 MMA, warps 2..9 epilogue
// XF (fused operand transform): warps 10.. transform the A patches, the warp after them loads the raw patches.  The
// 64-wide layers (one k-block per tile at Cin = 64: 4.1k cycles of MMA per patch) get 8 transform warps, the 128-wide 4.
constexpr int XF_SKEW = 128;                          // byte skew of the second patch plane (see the transform warps)
constexpr int TC_A_BYTES = 128 * 128;                 // 128 pixels x 64 fp16

template <int BN>
struct TcCfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 64) ? 4 : 3;
  // One TMEM partial-sum slot = 2*BN columns: [ hi*hi | hi*lo + lo*hi ].  tcgen05.mma (M=128, SS operands) has a floor of
  // ~107 cycles per instruction whatever N <= 128 is (tools/umma_rate.py: N=64 30 %, N=128 60 %, N=256 100 % of peak), so
  // the hi and lo weight tiles -- adjacent in shared memory -- are fed as ONE N = 2*BN operand: A_hi x [B_hi;B_lo] fills
  // both halves, A_lo x B_hi accumulates into the cross half.  2 MMAs per k-step instead of 3, and the small cross terms
  // never meet the large accumulator inside the truncating tensor-core adder.
  static constexpr int SLOT_COLS = 2 * BN;
  static constexpr int SLOTS = 512 / SLOT_COLS;            // 2 (BN=128) or 4 (BN=64)
  static constexpr int TMEM_COLS = 512;
  static constexpr int STG_BYTES = 8 * 4096;               // per-epilogue-warp 32x32-float transpose patches
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/ + STG_BYTES;
  // CTA-pair engine (cta_group::2, M = 256 over the two SMs of a TPC): each CTA stages its own A tile (hi, lo) and HALF of
  // every B operand: X = its half of [B_hi;B_lo] (rank 0: B_hi, rank 1: B_lo -> accumulator columns [0,BN) | [BN,2BN)) and
  // Y = its half of B_hi for the A_lo x B_hi pass (rank r: rows [r*BN/2, +BN/2))
  static constexpr int P_BX_BYTES = BN * 128;
  static constexpr int P_BY_BYTES = BN * 64;
  static constexpr int P_STAGE_BYTES = 2 * TC_A_BYTES + P_BX_BYTES + P_BY_BYTES;
  static constexpr int P_STAGES = (BN == 64) ? 4 : 3;
  static constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 512 + STG_BYTES;
  // halo engine: 16x8-pixel tiles; the (16+2)x(8+2) input patch of one 64-channel block is fetched ONCE (hi and lo
  // planes) and all 9 taps read it through row-shifted UMMA descriptors; weights stream through their own ring.
  static constexpr int H_A_PLANE = 23 * 1024;              // >= 18*10*128 B, 1024-aligned
  static constexpr int H_A_SLOT = 2 * H_A_PLANE;
  static constexpr int H_A_SLOTS = 2;
  static constexpr int H_B_SLOT = 2 * B_BYTES;
  static constexpr int H_B_SLOTS = (BN == 64) ? 6 : 3;
  static constexpr int H_SMEM_BYTES = H_A_SLOTS * H_A_SLOT + H_B_SLOTS * H_B_SLOT + 1024 + 512 + STG_BYTES;
  // halo engine in CTA-pair mode: B slots hold this CTA's halves (X | Y, see above)
  static constexpr int HP_B_SLOT = P_BX_BYTES + P_BY_BYTES;
  static constexpr int HP_B_SLOTS = (BN == 64) ? 8 : 4;
  static constexpr int HP_SMEM_BYTES = H_A_SLOTS * H_A_SLOT + HP_B_SLOTS * HP_B_SLOT + 1024 + 512 + STG_BYTES;
  // fused operand transform (XF, halo + pair only): the A patches arrive as the RAW fp32 activation (two 32-channel planes per
  // slot) and 4 / 8 transform warps apply GroupNorm-affine + SiLU + the hi/lo split in place before the MMAs read them;
  // one more A slot for the short-K (Cin = 64) layers, whose MMA time per patch is below TMA + transform latency
  static constexpr int XF_WARPS = (BN == 64) ? 8 : 4;
  static constexpr int XF_THREADS = TC_THREADS + 32 * XF_WARPS + 32;
  static constexpr int X_A_SLOTS = (BN == 64) ? 3 : 2;
  static constexpr int X_A_PLANE2 = H_A_PLANE + XF_SKEW;   // second plane of an XF slot (ends at 46720 <= H_A_SLOT)
  static constexpr int X_B_SLOTS = 4;
  static constexpr int X_SMEM_BYTES = X_A_SLOTS * H_A_SLOT + X_B_SLOTS * HP_B_SLOT + 1024 + 512 + STG_BYTES;
};

// Accumulation scheme (why the TMEM ring): tcgen05.mma adds into its fp32 accumulator with truncation, so a long
// K loop into one accumulator drifts by ~(#MMAs)*2^-25 relative (measured 2e-5 at K=4608 -- too much for the
// 1e-3 end-to-end bar).  Each TMEM slot therefore only receives `chunk` k-blocks (64 K-elements each), the small
// cross terms (lo*hi, hi*lo) are issued first while the slot is still tiny, and the epilogue warps fold every
// finished slot into fp32 registers with round-to-nearest adds while the tensor core fills the next slot.
// CPG > 0: the epilogue also emits GroupNorm(32) partial sums of the stored tile (CPG = Cout/32 channels per group).
// PAIR: two CTAs of a cluster (one TPC) work on two adjacent m-tiles with shared weights through tcgen05.mma.cta_group::2
// (M = 256): no single-CTA instruction floor (profiles/round1_umma_pair_probe.txt: N=128 at 100 % of the tensor rate vs
// 60 %) and each SM stages / reads only half of every B operand.  Rank 0 (leader) issues all MMAs; its `full` and `cempty`
// barriers collect the TMA bytes / epilogue arrivals of both CTAs; `empty` and `cfull` are signalled in both CTAs by
// multicast commits.
template <int BN, int CPG, bool HALO, bool PAIR, bool XF, bool GEN = false, bool K1 = false>
__global__ void __launch_bounds__(XF ? TcCfg<BN>::XF_THREADS : TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ CUtensorMap tmB_half, const TcParams p) {
  static_assert(!XF || (HALO && PAIR), "the fused operand transform exists for the halo + pair engine only");
  static_assert(!GEN || (XF && CPG == 0), "the generalised addressing exists for the fused-transform engine only");
  static_assert(!K1 || (XF && !GEN && CPG == 0), "K1 = fused transform of a 1x1 conv (patch = tile): its own instantiation");
  using Cfg = TcCfg<BN>;
  constexpr int A_SLOTS = XF ? Cfg::X_A_SLOTS : Cfg::H_A_SLOTS;
  constexpr int STAGES = PAIR ? Cfg::P_STAGES : Cfg::STAGES;
  constexpr int STAGE_BYTES = PAIR ? Cfg::P_STAGE_BYTES : Cfg::STAGE_BYTES;
  constexpr int TC_SLOTS = Cfg::SLOTS;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int first_tile = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // ring "full/empty": per-tap engine = STAGES k-block stages; halo engine = weight (B) slots.  "afull/aempty": halo A slots.
  constexpr int NRING = HALO ? (XF ? Cfg::X_B_SLOTS : (PAIR ? Cfg::HP_B_SLOTS : Cfg::H_B_SLOTS)) : STAGES;
  constexpr int RING_BYTES = HALO ? (PAIR ? Cfg::HP_B_SLOT : Cfg::H_B_SLOT) : STAGE_BYTES;
  uint8_t* ring_base = HALO ? smem + A_SLOTS * Cfg::H_A_SLOT : smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_base + NRING * RING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + NRING;
  uint64_t* cfull = bars + 2 * NRING;
  uint64_t* cempty = bars + 2 * NRING + TC_SLOTS;
  uint64_t* afull = bars + 2 * NRING + 2 * TC_SLOTS;
  uint64_t* aempty = afull + 3;
  uint64_t* araw = aempty + 3;                       // XF: raw patch landed (TMA -> transform warps), local to each CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(araw + 3);
  uint8_t* stage_buf = reinterpret_cast<uint8_t*>(bars) + 512;      // epilogue transpose patches (16-byte aligned)

  // Roles are numbered logically (0 TMA, 1 MMA, 2..9 epilogue, 10.. transform, then the patch loader) but sit on the warp
  // ids in REVERSE order: the sub-partition arbiter favours the highest warp id among its eligible warps, and the role that
  // must never wait for an issue slot is the MMA issuer, then the TMA producer, then the epilogue; the instruction-heavy
  // transform warps come last.  (TMEM lane quadrants follow the PHYSICAL warp id: lg below.)
  constexpr int NWARPS = (XF ? TcCfg<BN>::XF_THREADS : TC_THREADS) / 32;
  const int pwarp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int warp = NWARPS - 1 - pwarp;
  const int lane = threadIdx.x & 31;
  bool aborted = false;      // set when a barrier wait timed out anywhere on the device: leave the role loop (see mbar_wait)
  if (threadIdx.x == 0) {
    TC_STAMP(0);
#if CFB_TC_STAMPS
    if (p.dbg != nullptr && blockIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); p.dbg[14] = (long long)gt; }
#endif
  }

  if (warp == 0 && lane == 0) {
    for (int a = 0; a < 3; ++a) {
      mbar_init(smem_u32(afull + a), XF ? 2 * TcCfg<BN>::XF_WARPS : 1);      // XF: the transform warps of both CTAs arrive on the leader
      mbar_init(smem_u32(aempty + a), 1);
      mbar_init(smem_u32(araw + a), 1);
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
    for (int s = 0; s < NRING; ++s) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
    for (int a = 0; a < TC_SLOTS; ++a) {
      mbar_init(smem_u32(cfull + a), 1);
      mbar_init(smem_u32(cempty + a), PAIR ? 2 * TC_EPI_WARPS : TC_EPI_WARPS);   // pair: the epilogue warps of both CTAs
    }
    if constexpr (PAIR) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_half) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();      // both CTAs: barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above touched only this CTA's shared / tensor memory and the kernel parameters.
  // The next grid of the stream may start its own set-up now; this one waits here until the previous grid has completed.
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) TC_STAMP(1);

  // pair mode enumerates PAIRS of m-tiles: pm -> m-tiles (2pm, 2pm+1); Upsample keeps both CTAs on the same output
  // parity (shared weights): pm -> parity pm&3 of low-res tiles 2(pm>>2), 2(pm>>2)+1
  const int total_tiles = PAIR ? (p.m_tiles >> 1) * p.n_tiles : p.m_tiles * p.n_tiles;
  const int nk = p.taps * p.kblocks;
  auto mtile_of = [&](int pm) -> int {
    if constexpr (PAIR) return p.up4 ? (((((pm >> 2) << 1) + (int)rank) << 2) | (pm & 3)) : (pm * 2 + (int)rank);
    else return pm;
  };

  if (warp == 0) {
    // ============================ TMA producer (warp converged, one elected lane issues) ============================
    {
      int stage = 0;
      uint32_t phase = 0;
      int aslot = 0;
      uint32_t aphase = 0;
      const uint32_t rank0 = 0u;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int pm = tile / p.n_tiles, nt = tile - pm * p.n_tiles;
        const int mt = mtile_of(pm);
        const int mtl = p.up4 ? (mt >> 2) : mt;            // low-res tile; (mt & 3) = output parity (py,px)
        const int par_y = p.up4 ? ((mt & 3) >> 1) : 0, par_x = p.up4 ? (mt & 1) : 0;
        const int per_img = p.tiles_x * p.tiles_y;
        const int n = mtl / per_img;
        const int rem = mtl - n * per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int y0 = ty * p.BH * p.stride, x0 = tx * p.BW * p.stride;
        if constexpr (HALO) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            if constexpr (!XF) {
            mbar_wait<500>(smem_u32(aempty + aslot), aphase ^ 1, aborted); if (aborted) goto teardown;
            if (elect_one()) {
              const uint32_t sa = smem_u32(smem + aslot * Cfg::H_A_SLOT);
              if constexpr (PAIR) {
                if (rank == 0) mbar_expect_tx(smem_u32(afull + aslot), (uint32_t)(4 * p.PW * p.PH * 128));   // both CTAs' patches
                const uint32_t ab = map_to_cta(smem_u32(afull + aslot), rank0);
                tma_load_4d_pair(sa, &tmA_hi, ab, kb * 64, x0 - p.pad, y0 - p.pad, n);
                tma_load_4d_pair(sa + Cfg::H_A_PLANE, &tmA_lo, ab, kb * 64, x0 - p.pad, y0 - p.pad, n);
              } else {
                const uint32_t ab = smem_u32(afull + aslot);
                mbar_expect_tx(ab, (uint32_t)(2 * p.PW * p.PH * 128));
                tma_load_4d(sa, &tmA_hi, ab, kb * 64, x0 - p.pad, y0 - p.pad, n);
                tma_load_4d(sa + Cfg::H_A_PLANE, &tmA_lo, ab, kb * 64, x0 - p.pad, y0 - p.pad, n);
              }
            }
            __syncwarp();
            if (++aslot == A_SLOTS) { aslot = 0; aphase ^= 1; }
            }
            for (int tap = 0; tap < p.taps; ++tap) {
              const int btap = p.up4 ? (mt & 3) * 4 + tap : tap;      // Upsample: weight slice of this output parity
              mbar_wait<500>(smem_u32(empty + stage), phase ^ 1, aborted); if (aborted) goto teardown;
              if (elect_one()) {
                if (tile == first_tile && tap == 0 && kb == 0) TC_STAMP(2);
                const uint32_t sb = smem_u32(ring_base + stage * RING_BYTES);
                if constexpr (PAIR) {
                  if (rank == 0) mbar_expect_tx(smem_u32(full + stage), (uint32_t)(2 * Cfg::HP_B_SLOT));
                  const uint32_t fb = map_to_cta(smem_u32(full + stage), rank0);
                  const bool drop = p.fault && blockIdx.x == 0 && tile == first_tile && kb == 0 && tap == 0;   // injected fault
                  if (!drop) tma_load_3d_pair(sb, rank == 0 ? &tmB_hi : &tmB_lo, fb, kb * 64, nt * BN, btap);
                  tma_load_3d_pair(sb + Cfg::P_BX_BYTES, &tmB_half, fb, kb * 64, nt * BN + (int)rank * (BN / 2), btap);
                } else {
                  const uint32_t fb = smem_u32(full + stage);
                  mbar_expect_tx(fb, (uint32_t)Cfg::H_B_SLOT);
                  tma_load_3d(sb, &tmB_hi, fb, kb * 64, nt * BN, btap);
                  tma_load_3d(sb + Cfg::B_BYTES, &tmB_lo, fb, kb * 64, nt * BN, btap);
                }
              }
              __syncwarp();
              if (++stage == NRING) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          for (int tap = 0; tap < p.taps; ++tap) {
            // source offset of this tap and its weight slice: 3x3 -> (r,s)-pad; 2x2 parity conv -> (dy+py-1, dx+px-1)
            int r, s, btap = tap;
            if (p.up4) { r = (tap >> 1) + par_y; s = (tap & 1) + par_x; btap = (mt & 3) * 4 + tap; }
            else { r = (p.taps == 9) ? tap / 3 : 0; s = (p.taps == 9) ? tap - r * 3 : 0; }
            for (int kb = 0; kb < p.kblocks; ++kb) {
              mbar_wait<500>(smem_u32(empty + stage), phase ^ 1, aborted); if (aborted) goto teardown;
              if (elect_one()) {
                if (tile == first_tile && tap == 0 && kb == 0) TC_STAMP(2);
                const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                // multi-head batched GEMM: image index n = n_img * heads + head
                const int hh = p.heads > 1 ? n % p.heads : 0, nimg = p.heads > 1 ? n / p.heads : n;
                const int a_img = p.a_img_per_head ? n : nimg;
                const int ac = p.a_c0 + hh * p.a_c_head + kb * 64, bc = p.b_c0 + hh * p.b_c_head + kb * 64;
                const int brow = nt * BN + hh * p.b_r_head;
                const int b3 = p.b_batched ? nimg : btap;
                if constexpr (PAIR) {
                  // the leader's `full` barrier counts the bytes of both CTAs
                  if (rank == 0) mbar_expect_tx(smem_u32(full + stage), (uint32_t)(2 * Cfg::P_STAGE_BYTES));
                  const uint32_t fb = map_to_cta(smem_u32(full + stage), rank0);
                  tma_load_4d_pair(sa, &tmA_hi, fb, ac, x0 + s - p.pad, y0 + r - p.pad, a_img);
                  tma_load_4d_pair(sa + TC_A_BYTES, &tmA_lo, fb, ac, x0 + s - p.pad, y0 + r - p.pad, a_img);
                  tma_load_3d_pair(sa + 2 * TC_A_BYTES, rank == 0 ? &tmB_hi : &tmB_lo, fb, bc, brow, b3);
                  tma_load_3d_pair(sa + 2 * TC_A_BYTES + Cfg::P_BX_BYTES, &tmB_half, fb, bc, brow + (int)rank * (BN / 2), b3);
                } else {
                  const uint32_t fb = smem_u32(full + stage);
                  mbar_expect_tx(fb, (uint32_t)Cfg::STAGE_BYTES);
                  tma_load_4d(sa, &tmA_hi, fb, ac, x0 + s - p.pad, y0 + r - p.pad, a_img);
                  tma_load_4d(sa + TC_A_BYTES, &tmA_lo, fb, ac, x0 + s - p.pad, y0 + r - p.pad, a_img);
                  tma_load_3d(sa + 2 * TC_A_BYTES, &tmB_hi, fb, bc, brow, b3);
                  tma_load_3d(sa + 2 * TC_A_BYTES + Cfg::B_BYTES, &tmB_lo, fb, bc, brow, b3);
                }
              }
              __syncwarp();
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (warp converged, one elected lane issues) ============================
    {
      // kind::f16, A=B=F16 (0), D=F32 (1<<4), K-major A and B, N>>3 @17, M>>4 @24   (cute::UMMA::InstrDescriptor)
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);          // N = BN
      constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // N = 2*BN
      int stage = 0;
      uint32_t phase = 0;
      int slot = 0;
      uint32_t slot_phase = 0;
      if constexpr (HALO) {
        // A descriptors: rows of the tile are pixels (h, w) of a 16x8 patch; patch row h is one 8-row core-matrix
        // group that starts (h + r) * PW + s rows into the halo buffer => group stride SBO = PW*128 B and a start
        // address that is only 128-byte aligned.  The tensor core applies the 128B swizzle on absolute shared-memory
        // address bits (verified on B200 with tools/umma_probe.py: base_offset must stay 0), i.e. exactly the
        // pattern the TMA unit used when it wrote the buffer.
        const uint32_t a_desc_hi = (uint32_t)((p.PW * 128) >> 4) | (1u << 14) | (2u << 29);
        int aslot = 0;
        uint32_t aphase = 0;
        if constexpr (PAIR) {
          constexpr uint32_t pdesc2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
          constexpr uint32_t pdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
          if (rank == 0) {
            for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
              const int par = p.up4 ? ((tile / p.n_tiles) & 3) : 0;      // output parity of this (pair of) tile(s)
              const int par_y = par >> 1, par_x = par & 1;
              int it = 0;
              for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait_cl(smem_u32(afull + aslot), aphase, aborted); if (aborted) goto teardown;
                const uint32_t a_hi0 = smem_u32(smem + aslot * Cfg::H_A_SLOT), a_lo0 = a_hi0 + (XF ? Cfg::X_A_PLANE2 : Cfg::H_A_PLANE);
                for (int tap = 0; tap < p.taps; ++tap, ++it) {
                  int r = (p.taps == 9) ? tap / 3 : 0;
                  int sft = (p.taps == 9) ? tap - r * 3 : 0;
                  if (p.up4) { r = (tap >> 1) + par_y; sft = (tap & 1) + par_x; }
                  const bool first = (it % p.chunk) == 0;
                  if (first) mbar_wait_cl(smem_u32(cempty + slot), slot_phase ^ 1, aborted); if (aborted) goto teardown;
                  mbar_wait_cl(smem_u32(full + stage), phase, aborted); if (aborted) goto teardown;
                  tc_fence_after();
                  if (elect_one()) {
                    if (tile == first_tile && it == 0) TC_STAMP(3);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(slot * Cfg::SLOT_COLS);
                    const uint32_t aoff = (uint32_t)((r * p.PW + sft) * 128);
                    const uint32_t ah = desc_lo(a_hi0 + aoff), al = desc_lo(a_lo0 + aoff);
                    const uint32_t sb = smem_u32(ring_base + stage * RING_BYTES);
                    const uint32_t bx = desc_lo(sb), by = desc_lo(sb + Cfg::P_BX_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      tc_mma_f16_pair_w(d_tmem, ah + 2 * k, a_desc_hi, bx + 2 * k, DESC_HI_SW128, pdesc2, (!first || k > 0) ? 1u : 0u);
                      tc_mma_f16_pair_w(d_tmem + BN, al + 2 * k, a_desc_hi, by + 2 * k, DESC_HI_SW128, pdesc, 1u);
                    }
                    tc_commit_pair(smem_u32(empty + stage), (uint16_t)3);
                    if ((it % p.chunk) == p.chunk - 1 || it == nk - 1) tc_commit_pair(smem_u32(cfull + slot), (uint16_t)3);
                    if (tap == p.taps - 1) tc_commit_pair(smem_u32(aempty + aslot), (uint16_t)3);
                    if (tile == first_tile && it == nk - 1) TC_STAMP(4);
                  }
                  __syncwarp();
                  if (++stage == NRING) { stage = 0; phase ^= 1; }
                  if ((it % p.chunk) == p.chunk - 1 || it == nk - 1) {
                    if (++slot == TC_SLOTS) { slot = 0; slot_phase ^= 1; }
                  }
                }
                if (++aslot == A_SLOTS) { aslot = 0; aphase ^= 1; }
              }
            }
          }
        } else {
        for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
          const int par = p.up4 ? ((tile / p.n_tiles) & 3) : 0;
          const int par_y = par >> 1, par_x = par & 1;
          int it = 0;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(smem_u32(afull + aslot), aphase, aborted); if (aborted) goto teardown;
            const uint32_t a_hi0 = smem_u32(smem + aslot * Cfg::H_A_SLOT), a_lo0 = a_hi0 + Cfg::H_A_PLANE;
            for (int tap = 0; tap < p.taps; ++tap, ++it) {
              int r = (p.taps == 9) ? tap / 3 : 0;
              int sft = (p.taps == 9) ? tap - r * 3 : 0;
              if (p.up4) { r = (tap >> 1) + par_y; sft = (tap & 1) + par_x; }
              const bool first = (it % p.chunk) == 0;
              if (first) mbar_wait(smem_u32(cempty + slot), slot_phase ^ 1, aborted); if (aborted) goto teardown;
              mbar_wait(smem_u32(full + stage), phase, aborted); if (aborted) goto teardown;
              tc_fence_after();
              if (elect_one()) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(slot * Cfg::SLOT_COLS);
                const uint32_t aoff = (uint32_t)((r * p.PW + sft) * 128);
                const uint32_t ah = desc_lo(a_hi0 + aoff), al = desc_lo(a_lo0 + aoff);
                const uint32_t bh = desc_lo(smem_u32(ring_base + stage * RING_BYTES));   // [B_hi ; B_lo] contiguous rows
                // 64-wide k-block = 4 x UMMA_K(16): +32 B (= +2 in the >>4 address field) inside the swizzle atom.
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  tc_mma_f16_w(d_tmem, ah + 2 * k, a_desc_hi, bh + 2 * k, DESC_HI_SW128, idesc2, (!first || k > 0) ? 1u : 0u);
                  tc_mma_f16_w(d_tmem + BN, al + 2 * k, a_desc_hi, bh + 2 * k, DESC_HI_SW128, idesc, 1u);
                }
                tc_commit(smem_u32(empty + stage));
                if ((it % p.chunk) == p.chunk - 1 || it == nk - 1) tc_commit(smem_u32(cfull + slot));
                if (tap == p.taps - 1) tc_commit(smem_u32(aempty + aslot));   // every tap of this block has read the patch
              }
              __syncwarp();
              if (++stage == NRING) { stage = 0; phase ^= 1; }
              if ((it % p.chunk) == p.chunk - 1 || it == nk - 1) {
                if (++slot == TC_SLOTS) { slot = 0; slot_phase ^= 1; }
              }
            }
            if (++aslot == A_SLOTS) { aslot = 0; aphase ^= 1; }
          }
        }
        }
      } else if constexpr (PAIR) {
        // M = 256 over the pair; per CTA half operands: X (this CTA's half of [B_hi;B_lo]) and Y (its half of B_hi)
        constexpr uint32_t pdesc2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // N = 2*BN
        constexpr uint32_t pdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);          // N = BN
        if (rank == 0) {
          for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
            for (int it0 = 0; it0 < nk; it0 += p.chunk) {
              mbar_wait_cl(smem_u32(cempty + slot), slot_phase ^ 1, aborted); if (aborted) goto teardown;
              const int it1 = (it0 + p.chunk < nk) ? it0 + p.chunk : nk;
              for (int it = it0; it < it1; ++it) {
                mbar_wait_cl(smem_u32(full + stage), phase, aborted); if (aborted) goto teardown;
                tc_fence_after();
                if (elect_one()) {
                  if (tile == first_tile && it == 0) TC_STAMP(3);
                  const uint32_t d_tmem = tmem_base + (uint32_t)(slot * Cfg::SLOT_COLS);
                  const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                  const uint32_t ah = desc_lo(sa), al = desc_lo(sa + TC_A_BYTES);
                  const uint32_t bx = desc_lo(sa + 2 * TC_A_BYTES), by = desc_lo(sa + 2 * TC_A_BYTES + Cfg::P_BX_BYTES);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    tc_mma_f16_pair_w(d_tmem, ah + 2 * k, DESC_HI_SW128, bx + 2 * k, DESC_HI_SW128, pdesc2, (it > it0 || k > 0) ? 1u : 0u);
                    tc_mma_f16_pair_w(d_tmem + BN, al + 2 * k, DESC_HI_SW128, by + 2 * k, DESC_HI_SW128, pdesc, 1u);
                  }
                  tc_commit_pair(smem_u32(empty + stage), (uint16_t)3);             // stage reusable in BOTH CTAs
                  if (it == it1 - 1) tc_commit_pair(smem_u32(cfull + slot), (uint16_t)3);   // both epilogues fold their rows
                  if (tile == first_tile && it == nk - 1) TC_STAMP(4);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
              }
              if (++slot == TC_SLOTS) { slot = 0; slot_phase ^= 1; }
            }
          }
        }
      } else {
        for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
          for (int it0 = 0; it0 < nk; it0 += p.chunk) {
            mbar_wait(smem_u32(cempty + slot), slot_phase ^ 1, aborted); if (aborted) goto teardown;
            const int it1 = (it0 + p.chunk < nk) ? it0 + p.chunk : nk;
            for (int it = it0; it < it1; ++it) {
              mbar_wait(smem_u32(full + stage), phase, aborted); if (aborted) goto teardown;
              tc_fence_after();
              if (elect_one()) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(slot * Cfg::SLOT_COLS);
                const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t ah = desc_lo(sa), al = desc_lo(sa + TC_A_BYTES), bh = desc_lo(sa + 2 * TC_A_BYTES);   // [B_hi;B_lo]
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  tc_mma_f16_w(d_tmem, ah + 2 * k, DESC_HI_SW128, bh + 2 * k, DESC_HI_SW128, idesc2, (it > it0 || k > 0) ? 1u : 0u);
                  tc_mma_f16_w(d_tmem + BN, al + 2 * k, DESC_HI_SW128, bh + 2 * k, DESC_HI_SW128, idesc, 1u);
                }
                tc_commit(smem_u32(empty + stage));           // smem slot reusable once these MMAs have read it
                if (it == it1 - 1) tc_commit(smem_u32(cfull + slot));   // partial sum complete -> epilogue warps fold it
              }
              __syncwarp();
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++slot == TC_SLOTS) { slot = 0; slot_phase ^= 1; }
          }
        }
      }
    }
  } else if (XF && warp == 2 + TC_EPI_WARPS + TcCfg<BN>::XF_WARPS) {
    // ============================ XF: A-patch loader (own warp: patches must be requested a whole patch ahead) ==========
    // The patch of one 64-channel block arrives as RAW fp32 NHWC values of the producing conv's output: two TMA boxes of
    // 32 channels (128 B rows, 128B swizzle) into the two planes of the slot.  tmA_hi / tmA_lo are the fp32 tensor maps of
    // source 0 / source 1: the k-blocks [0, a_split) come from source 0, the rest from source 1 -- torch.cat([enc, dec])
    // of Fuse_sft_block (codeformer_arch.py:152) never exists in memory.
    if constexpr (XF) {
      int aslot = 0;
      uint32_t aphase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int mt0 = mtile_of(tile / p.n_tiles);
        const int mt = p.up4 ? (mt0 >> 2) : mt0;           // Upsample: the four output parities read the same low-res patch
        const int per_img = p.tiles_x * p.tiles_y;
        const int n = mt / per_img;
        const int rem = mt - n * per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int y0 = ty * p.BH, x0 = tx * p.BW;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait<500>(smem_u32(aempty + aslot), aphase ^ 1, aborted); if (aborted) goto teardown;       // every MMA that read this slot has completed
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + aslot * Cfg::H_A_SLOT);
            const uint32_t rb = smem_u32(araw + aslot);
            const bool src0 = kb < p.a_split;
            const CUtensorMap* src = src0 ? &tmA_hi : &tmA_lo;
            const int c0 = (src0 ? kb : kb - p.a_split) * 64;
            if (tile == first_tile && kb == 0) TC_STAMP(10);
            mbar_expect_tx(rb, (uint32_t)(2 * p.PW * p.PH * 128));
            tma_load_4d(sa, src, rb, c0, x0 - p.pad, y0 - p.pad, n);
            tma_load_4d(sa + Cfg::X_A_PLANE2, src, rb, c0 + 32, x0 - p.pad, y0 - p.pad, n);
          }
          __syncwarp();
          if (++aslot == A_SLOTS) { aslot = 0; aphase ^= 1; }
        }
      }
    }
  } else if (XF && warp >= 2 + TC_EPI_WARPS) {
    // ============================ XF: operand transform (warps 10..) ============================
    // raw fp32 patch (zero outside the image: TMA out-of-bounds fill) -> y = act(x * scale[n,c] + shift[n,c]) (GroupNorm folded
    // into scale/shift, vqgan_arch.py:14-20,153-160) -> fp16 hi = rn(y), lo = rn(y - hi), written back IN PLACE in the
    // 128B-swizzled K-major layout the MMA descriptors read: plane 0 (raw channels 0..31 of the block) becomes the hi plane
    // (64 channels x fp16), plane 1 (raw channels 32..63) the lo plane.  A patch row (one pixel, 256 B raw) is handled by 8
    // lanes of ONE warp -- lane j owns channels 8j..8j+7 -- and every lane loads before any lane stores (__syncwarp), which
    // is what makes the in-place rewrite safe.  Plane 1 is skewed by 128 B so that, with the swizzle being a function of
    // absolute shared-memory address bits, the 16-byte chunks {0,2,4,6} of plane-0 lanes and plane-1 lanes of the same row
    // fall on different banks: each warp-wide LDS.128 / STS.128 is 4 conflict-free wavefronts.  Pixels outside the image
    // stay exactly zero (the conv pads the NORMALISED tensor).  fence.proxy.async publishes the writes to the tensor core.
    if constexpr (XF) {
      constexpr int XFW = TcCfg<BN>::XF_WARPS;
      constexpr int RPP = 4 * XFW;                                  // patch rows per pass (8 lanes per row)
      constexpr int XF_PW = 10, XF_PH = 18, XF_ROWS = XF_PW * XF_PH;   // halo patch of an 8 x 16 tile and a 3 x 3 filter
      constexpr int NPASS = (XF_ROWS + RPP - 1) / RPP;
      constexpr int NPASS1 = (128 + RPP - 1) / RPP;                     // 1x1 convs: the patch is the 8 x 16 tile itself
      const int t = (warp - (2 + TC_EPI_WARPS)) * 32 + lane;
      const int j = t & 7, rsub = t >> 3;
      const int pl = j >> 2, c0 = 2 * (j & 3);                      // source plane and first 16-byte chunk of this lane's 8 channels
      const uint32_t afull_leader = map_to_cta(smem_u32(afull), 0u);
      const int mode = p.in_scale ? (p.in_act == IN_SILU ? 2 : 1) : 0;   // 2: affine + SiLU, 1: affine, 0: raw split
      const int Hin = GEN ? p.Hin : p.tiles_y * p.BH, Win = GEN ? p.Win : p.tiles_x * p.BW;
      const int Cin = p.kblocks * 64;
      float amax = 0.f;
      int aslot = 0;
      uint32_t aphase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int mt0 = mtile_of(tile / p.n_tiles);
        const int mt = p.up4 ? (mt0 >> 2) : mt0;
        const int per_img = p.tiles_x * p.tiles_y;
        const int n = mt / per_img;
        const int rem = mt - n * per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int y0 = ty * p.BH - p.pad, x0 = tx * p.BW - p.pad;
        const bool border = !K1 && (y0 < 0 || x0 < 0 || y0 + XF_PH > Hin || x0 + XF_PW > Win);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          float sc[8], sh[8];
          if (mode) {
            const int nn = GEN ? min(n, p.N - 1) : n;       // GEN: the tile count is padded to an even number (dummy tile)
            const float* sp = p.in_scale + (int64_t)nn * Cin + kb * 64 + j * 8;
            const float* hp = p.in_shift + (int64_t)nn * Cin + kb * 64 + j * 8;
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp)), s1 = __ldg(reinterpret_cast<const float4*>(sp + 4));
            const float4 h0 = __ldg(reinterpret_cast<const float4*>(hp)), h1 = __ldg(reinterpret_cast<const float4*>(hp + 4));
            sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
            sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) { sc[k] = 1.f; sh[k] = 0.f; }
          }
          mbar_wait<250>(smem_u32(araw + aslot), aphase, aborted); if (aborted) goto teardown;
          if (t == 0 && tile == first_tile && kb == 0) TC_STAMP(11);
          if (t == 0 && tile == first_tile + tile_step && kb == 0) TC_STAMP(19);
          const uint32_t base0 = smem_u32(smem + aslot * Cfg::H_A_SLOT);
          const uint32_t src_base = base0 + (pl ? (uint32_t)Cfg::X_A_PLANE2 : 0u);
          const uint32_t lo_base = base0 + (uint32_t)Cfg::X_A_PLANE2;
          if constexpr (K1) {
            if (mode == 2) xf_patch<2, RPP, NPASS1, 8, 128>(src_base, base0, lo_base, c0, j, rsub, sc, sh, border, y0, x0, Hin, Win, amax);
            else if (mode == 1) xf_patch<1, RPP, NPASS1, 8, 128>(src_base, base0, lo_base, c0, j, rsub, sc, sh, border, y0, x0, Hin, Win, amax);
            else xf_patch<0, RPP, NPASS1, 8, 128>(src_base, base0, lo_base, c0, j, rsub, sc, sh, border, y0, x0, Hin, Win, amax);
          } else {
            if (mode == 2) xf_patch<2, RPP, NPASS, XF_PW, XF_ROWS>(src_base, base0, lo_base, c0, j, rsub, sc, sh, border, y0, x0, Hin, Win, amax);
            else if (mode == 1) xf_patch<1, RPP, NPASS, XF_PW, XF_ROWS>(src_base, base0, lo_base, c0, j, rsub, sc, sh, border, y0, x0, Hin, Win, amax);
            else xf_patch<0, RPP, NPASS, XF_PW, XF_ROWS>(src_base, base0, lo_base, c0, j, rsub, sc, sh, border, y0, x0, Hin, Win, amax);
          }
          if constexpr (GEN) {
            if (p.pad_mode && border) {
              // ReflectionPad2d / replicate padding: an out-of-image pixel of the patch equals an in-image pixel of the SAME
              // patch (index -1 -> 1 or 0, index H -> H-2 or H-1), so after every warp has written its rows the outside rows
              // are copied from their source rows (transformed values: the per-channel affine is position independent)
              asm volatile("bar.sync 1, %0;" ::"r"(32 * XFW) : "memory");
#pragma unroll 1
              for (int r = rsub; r < XF_ROWS; r += RPP) {
                const int py = r / XF_PW, px = r - py * XF_PW;
                const int gy = y0 + py, gx = x0 + px;
                if ((unsigned)gy < (unsigned)Hin && (unsigned)gx < (unsigned)Win) continue;
                int sy = gy, sx = gx;
                if (p.pad_mode == 1) {
                  sy = sy < 0 ? -sy : (sy >= Hin ? 2 * Hin - 2 - sy : sy);
                  sx = sx < 0 ? -sx : (sx >= Win ? 2 * Win - 2 - sx : sx);
                }
                sy = min(max(sy, max(y0, 0)), min(Hin, y0 + XF_PH) - 1);      // inside the image AND inside this patch
                sx = min(max(sx, max(x0, 0)), min(Win, x0 + XF_PW) - 1);
                const int rs = (sy - y0) * XF_PW + (sx - x0);
                const uint32_t hs = base0 + (uint32_t)rs * 128u, hd = base0 + (uint32_t)r * 128u;
                const uint32_t ls = lo_base + (uint32_t)rs * 128u, ld = lo_base + (uint32_t)r * 128u;
                const uint4 hv = lds128(hs + ((((uint32_t)j) ^ ((hs >> 7) & 7u)) << 4));
                const uint4 lv = lds128(ls + ((((uint32_t)j) ^ ((ls >> 7) & 7u)) << 4));
                sts128(hd + ((((uint32_t)j) ^ ((hd >> 7) & 7u)) << 4), hv);
                sts128(ld + ((((uint32_t)j) ^ ((ld >> 7) & 7u)) << 4), lv);
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(afull_leader + (uint32_t)(aslot * 8));
          if (t == 0 && tile == first_tile && kb == 0) TC_STAMP(12);
          if (t == 0 && tile == first_tile + tile_step && kb == 0) TC_STAMP(20);
          if (++aslot == A_SLOTS) { aslot = 0; aphase ^= 1; }
        }
      }
      if (amax > 65504.f) report_overflow();      // an operand left the fp16 range: the host turns the status word into an error
    }
  } else {
    // ============================ epilogue (warps 2..9) ============================
    constexpr int HC = BN / 2;               // columns owned by this thread
    const int lg = pwarp & 3;                // TMEM lane quadrant this warp may access: lanes [32*lg, 32*lg+32) (physical warp id)
    const int half = (warp - 2) >> 2;        // which half of the tile's columns (logical warps e and e+4 share a quadrant)
    const int row = lg * 32 + lane;          // pixel row of the tile
    const int cbase = half * HC;
    const float wsi = __ldg(p.wscale_inv);
    // tile geometry: 8 x 16 pixels on the halo engine (compile-time), else BW = 2^bw_shift columns (or any BW: division)
    const int BW = HALO ? 8 : p.BW, BH = HALO ? 16 : p.BH;
    const int bw_shift = HALO ? 3 : ((BW & (BW - 1)) == 0 ? __ffs(BW) - 1 : -1);   // log2(BW) when BW is a power of two
    // element strides of one tile row / column in `out` (Upsample tiles write every second pixel of the output)
    const int64_t SW = (int64_t)(p.up4 ? 2 : 1) * p.Cout, SH = SW * p.Wo;
    int slot = 0;
    uint32_t slot_phase = 0;
    float omax = 0.f;                        // largest magnitude emitted into fp16 operand planes (range guard)
    const uint32_t cempty_leader = PAIR ? map_to_cta(smem_u32(cempty), 0u) : 0u;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      const int pm = tile / p.n_tiles, nt = tile - pm * p.n_tiles;
      const int mt = mtile_of(pm);
      const int mtl = p.up4 ? (mt >> 2) : mt;
      const int per_img = p.tiles_x * p.tiles_y;
      const int nb = mtl / per_img;
      const int rem = mtl - nb * per_img;
      // multi-head batched GEMM writing [n][token][heads * o_c_head]: head h of image n owns its column slice
      const bool hsplit = !HALO && p.heads > 1 && !p.out_per_head;       // batched-GEMM kernels only (per-tap engine)
      const int n = hsplit ? nb / p.heads : nb;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int h = bw_shift >= 0 ? (row >> bw_shift) : row / BW, w = row - h * BW;
      int oy = ty * BH + h, ox = tx * BW + w;
      if (p.up4) { oy = 2 * oy + ((mt & 3) >> 1); ox = 2 * ox + (mt & 1); }   // this tile writes one output parity
      const int64_t pix = ((int64_t)n * p.Ho + oy) * p.Wo + ox;
      // first pixel of the tile in `out` (elements); the (row, chunk) items of the store loop are hh * SH + ww * SW away
      const int64_t off_tile = (((int64_t)n * p.Ho + (p.up4 ? 2 * ty * BH + ((mt & 3) >> 1) : ty * BH)) * p.Wo +
                                (p.up4 ? 2 * tx * BW + (mt & 1) : tx * BW)) * p.Cout;
      const int col0 = nt * BN + cbase + (hsplit ? (nb % p.heads) * p.o_c_head : 0);
      const int64_t off0 = pix * p.Cout + col0;
      // pull this thread's residual / SFT row slices towards L2 now: they are consumed only after the whole K loop
      if (!GEN && p.residual) {
#pragma unroll
        for (int j = 0; j < HC; j += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + off0 + j));
      }
      if (p.sft_dec) {
#pragma unroll
        for (int j = 0; j < HC; j += 32) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sft_dec + off0 + j));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sft_scale + off0 + j));
        }
      }
      float acc[HC];
#pragma unroll
      for (int j = 0; j < HC; ++j) acc[j] = 0.f;
      for (int it0 = 0; it0 < nk; it0 += p.chunk) {
        mbar_wait<250>(smem_u32(cfull + slot), slot_phase, aborted); if (aborted) goto teardown;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(slot * Cfg::SLOT_COLS + cbase);
        if constexpr (XF) {
          // the transform variants run 15 / 19 warps per CTA (128 / 96 registers per thread): 16-column TMEM chunks keep the
          // fold inside the register budget (32-column chunks spilled the accumulators)
#pragma unroll
          for (int c0 = 0; c0 < HC; c0 += 16) {
            uint32_t r0[16], r1[16];
            tmem_ld16(taddr + c0, r0);
            tmem_ld16(taddr + BN + c0, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
          }
        } else {
#pragma unroll
          for (int c0 = 0; c0 < HC; c0 += 32) {          // main half at +0, cross half at +BN; round-to-nearest adds
            uint32_t r0[32], r1[32];
            tmem_ld32(taddr + c0, r0);
            tmem_ld32(taddr + BN + c0, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster_relaxed(cempty_leader + (uint32_t)(slot * 8));
          else mbar_arrive(smem_u32(cempty + slot));
        }
        if (++slot == TC_SLOTS) { slot = 0; slot_phase ^= 1; }
      }
      if (warp == 2 && lane == 0 && tile == first_tile) TC_STAMP(7);
      if (warp == 2 && lane == 0 && tile == first_tile + tile_step) TC_STAMP(17);
      // ---- finalize this tile: scale, bias, residual, activation, SFT, store (fp32 NHWC), GroupNorm partials.
      // A TMEM lane owns a pixel ROW, so storing straight from registers would touch 32 different 128-byte lines per
      // instruction.  Each warp instead transposes 32x32-float blocks through a private 4 KB XOR-swizzled smem patch:
      // afterwards lane l holds the 16-byte chunk (l & 7) of row (l >> 3) + 4*it, i.e. 8 lanes cover one full 128-byte
      // line and every global access (residual / SFT loads, the store) is a fully used line.
      if (!HALO && !GEN && p.vq_cand) {
        // VectorQuantizer.forward (vqgan_arch.py:40-46): this thread owns token row `pix` and HC codes; d = (|z|^2 + |e|^2) - 2 z.e
        // in the reference's operation order, first minimum of the slice (ascending index, strict <); no staging, no store of
        // the [tokens, codes] matrix.  The candidates of a token (2 per n-tile) are reduced by vq_select_cand.
        const float z2 = __ldg(p.vq_z2 + pix);
        float best = INFINITY, dsum = 0.f;
        int bi = col0;
#pragma unroll
        for (int j = 0; j < HC; ++j) {
          const float d = (z2 + __ldg(p.vq_e2 + col0 + j)) - 2.f * (acc[j] * wsi);
          dsum += d;
          if (d < best) { best = d; bi = col0 + j; }
        }
        p.vq_cand[pix * (2 * p.n_tiles) + nt * 2 + half] = make_float2(best, __int_as_float(bi));
        double ds = (double)dsum;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
        if (lane == 0) p.vq_dpart[((int64_t)mt * p.n_tiles + nt) * 8 + (warp - 2)] = ds;
        continue;
      }
      const uint32_t stg = smem_u32(stage_buf) + (uint32_t)(warp - 2) * 4096u;     // 32 rows x 8 chunks of 16 B
      const uint32_t stg_w = stg + (uint32_t)lane * 128u;                          // this lane's row (write side)
      const int cch = lane & 7, rsub = lane >> 3;
      if constexpr (GEN) {
        // generalised placement (ParseNet / RRDBNet): ragged tiles (only pixels inside the true image are stored), destination
        // with its own channel pitch / offset (dense-block buffers), optional even-position subsampling (stride 2), and
        //   out = act(conv * 2^-k + bias + residual) * post_scale + residual2
        const int Hs = p.sub ? (p.Ho >> 1) : p.Ho, Ws = p.sub ? (p.Wo >> 1) : p.Wo;
#pragma unroll
        for (int q = 0; q < HC; q += 32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128f(stg_w + (uint32_t)((j ^ (lane & 7)) << 4), acc[q + 4 * j] * wsi, acc[q + 4 * j + 1] * wsi, acc[q + 4 * j + 2] * wsi,
                    acc[q + 4 * j + 3] * wsi);
          __syncwarp();
          const int colq = col0 + q + cch * 4;
          const bool col_ok = colq < p.cout_valid;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + colq));
          int64_t pixs[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int trow = lg * 32 + it * 4 + rsub;
            const int hh = trow / BW, ww = trow - hh * BW;
            int oy2 = ty * BH + hh, ox2 = tx * BW + ww;
            if (p.up4) { oy2 = 2 * oy2 + ((mt & 3) >> 1); ox2 = 2 * ox2 + (mt & 1); }
            bool ok = col_ok && n < p.N && oy2 < p.Ho && ox2 < p.Wo;
            if (p.sub) { ok = ok && (((oy2 | ox2) & 1) == 0); oy2 >>= 1; ox2 >>= 1; }
            pixs[it] = ok ? ((int64_t)n * Hs + oy2) * Ws + ox2 : (int64_t)-1;
          }
          float4 rres[8];
#pragma unroll
          for (int it = 0; it < 8; ++it)
            rres[it] = (p.residual && pixs[it] >= 0) ? __ldg(reinterpret_cast<const float4*>(p.residual + pixs[it] * p.res_pitch + colq))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            float4 v = lds128f(stg + (uint32_t)(r * 128 + ((cch ^ (r & 7)) << 4)));
            v.x += bv.x + rres[it].x; v.y += bv.y + rres[it].y; v.z += bv.z + rres[it].z; v.w += bv.w + rres[it].w;
            if (p.out_act == OUT_LRELU) {
              v.x = v.x > 0.f ? v.x : 0.2f * v.x; v.y = v.y > 0.f ? v.y : 0.2f * v.y;
              v.z = v.z > 0.f ? v.z : 0.2f * v.z; v.w = v.w > 0.f ? v.w : 0.2f * v.w;
            }
            if (pixs[it] >= 0) {
              if (p.residual2) {
                const float4 r2 = __ldg(reinterpret_cast<const float4*>(p.residual2 + pixs[it] * p.res2_pitch + colq));
                v.x = fmaf(v.x, p.post_scale, r2.x); v.y = fmaf(v.y, p.post_scale, r2.y);
                v.z = fmaf(v.z, p.post_scale, r2.z); v.w = fmaf(v.w, p.post_scale, r2.w);
              }
              *reinterpret_cast<float4*>(p.out + pixs[it] * p.out_pitch + p.out_c0 + colq) = v;
            }
          }
          __syncwarp();
        }
      } else {
#pragma unroll
      for (int q = 0; q < HC; q += 32) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128f(stg_w + (uint32_t)((j ^ (lane & 7)) << 4), acc[q + 4 * j] * wsi, acc[q + 4 * j + 1] * wsi, acc[q + 4 * j + 2] * wsi,
                  acc[q + 4 * j + 3] * wsi);
        __syncwarp();
        const int colq = col0 + q + cch * 4;                  // first of this lane's 4 channels
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + colq));
        // global offsets of the (row, chunk) items of this lane, then their residual loads in flight at once: all 8 rows, or two
        // batches of 4 in the register-capped transform variants
        constexpr int RB = XF ? 4 : 8;
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
        for (int ib = 0; ib < 8; ib += RB) {
        int64_t offs[RB];
#pragma unroll
        for (int k = 0; k < RB; ++k) {
          const int it = ib + k;
          const int trow = lg * 32 + it * 4 + rsub;
          const int hh = bw_shift >= 0 ? (trow >> bw_shift) : trow / BW, ww = trow - hh * BW;
          offs[k] = off_tile + colq + hh * SH + ww * SW;
        }
        float4 rres[RB];
#pragma unroll
        for (int k = 0; k < RB; ++k)
          rres[k] = p.residual ? __ldg(reinterpret_cast<const float4*>(p.residual + offs[k])) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < RB; ++k) {
          const int it = ib + k;
          const int r = it * 4 + rsub;                        // row within this warp's 32-row quadrant
          float4 v = lds128f(stg + (uint32_t)(r * 128 + ((cch ^ (r & 7)) << 4)));
          const int64_t off = offs[k];
          v.x += bv.x + rres[k].x; v.y += bv.y + rres[k].y; v.z += bv.z + rres[k].z; v.w += bv.w + rres[k].w;
          if (p.out_act == OUT_LRELU) {
            v.x = v.x > 0.f ? v.x : 0.2f * v.x; v.y = v.y > 0.f ? v.y : 0.2f * v.y;
            v.z = v.z > 0.f ? v.z : 0.2f * v.z; v.w = v.w > 0.f ? v.w : 0.2f * v.w;
          } else if (p.out_act == OUT_GELU) {
            v.x = 0.5f * v.x * (1.f + erff(v.x * 0.70710678118654752440f));
            v.y = 0.5f * v.y * (1.f + erff(v.y * 0.70710678118654752440f));
            v.z = 0.5f * v.z * (1.f + erff(v.z * 0.70710678118654752440f));
            v.w = 0.5f * v.w * (1.f + erff(v.w * 0.70710678118654752440f));
          }
          if (p.sft_dec) {
            const float4 d = __ldg(reinterpret_cast<const float4*>(p.sft_dec + off));
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.sft_scale + off));
            v.x = d.x + p.sft_w * (d.x * sc.x + v.x); v.y = d.y + p.sft_w * (d.y * sc.y + v.y);
            v.z = d.z + p.sft_w * (d.z * sc.z + v.z); v.w = d.w + p.sft_w * (d.w * sc.w + v.w);
          }
          if (p.out) *reinterpret_cast<float4*>(p.out + off) = v;      // null: only the operand planes are consumed
          if (p.pl_hi) {
            omax = fmaxf(omax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
            uint2 ph, pl;
            ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
            pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
            *reinterpret_cast<uint2*>(p.pl_hi + off) = ph;
            *reinterpret_cast<uint2*>(p.pl_lo + off) = pl;
          }
          if constexpr (CPG == 2) {
            s0 += v.x + v.y; q0 += fmaf(v.x, v.x, v.y * v.y);
            s1 += v.z + v.w; q1 += fmaf(v.z, v.z, v.w * v.w);
          } else if constexpr (CPG >= 4) {
            s0 += (v.x + v.y) + (v.z + v.w);
            q0 += fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
          }
        }
        }
        if constexpr (CPG > 0) {
          // GroupNorm partial sums of the values just stored: reduce over the 4 row-lanes (xor 8, 16) and, for groups
          // wider than one chunk, over the chunk-lanes of the group; fixed order => deterministic
          s0 += __shfl_xor_sync(0xffffffffu, s0, 8); q0 += __shfl_xor_sync(0xffffffffu, q0, 8);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 16); q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
          if constexpr (CPG == 2) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, 8); q1 += __shfl_xor_sync(0xffffffffu, q1, 8);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
          }
          if constexpr (CPG >= 8) { s0 += __shfl_xor_sync(0xffffffffu, s0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 1); }
          if constexpr (CPG >= 16) { s0 += __shfl_xor_sync(0xffffffffu, s0, 2); q0 += __shfl_xor_sync(0xffffffffu, q0, 2); }
          constexpr int CL = (CPG >= 4) ? CPG / 4 : 1;         // chunk-lanes per group
          if (rsub == 0 && (cch & (CL - 1)) == 0) {
            float* gp = p.gn_part + ((int64_t)mt * 4 + lg) * 64;
            if constexpr (CPG == 2) {
              *reinterpret_cast<float4*>(gp + (colq / 2) * 2) = make_float4(s0, q0, s1, q1);
            } else {
              *reinterpret_cast<float2*>(gp + (colq / CPG) * 2) = make_float2(s0, q0);
            }
          }
        }
        __syncwarp();
      }
      }
      if (warp == 2 && lane == 0 && tile == first_tile) TC_STAMP(16);
      if (warp == 2 && lane == 0 && tile == first_tile + tile_step) TC_STAMP(18);
    }
    if (omax > 65504.f) report_overflow();   // a value left the fp16 range of the operand planes: reported, never silent
    if (warp == 2 && lane == 0) TC_STAMP(13);
  }

teardown:
  if (aborted) {             // error path only: let the bulk copies / MMAs that are still in flight finish before the CTA's
    const long long t0 = clock64();          // shared and tensor memory are handed back
    while (clock64() - t0 < 400000) {}
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_STAMP(8);
  if constexpr (PAIR) cluster_sync_all();      // the peer may still read this CTA's operands / signal its barriers
  if (threadIdx.x == 0) {
    TC_STAMP(9);
#if CFB_TC_STAMPS
    if (p.dbg != nullptr && blockIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); p.dbg[15] = (long long)gt; }
#endif
  }
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------
// asynchronous status word: binding of this device's symbols (called once per device by runtime.cu)
// ------------------------------------------------------------------------------------------------------
int tc_bind_status_word(unsigned* host_mapped_dev_ptr, long long wait_limit_cycles) {
  const unsigned zero = 0;
  CFB_CUDA(cudaMemcpyToSymbol(g_status_host, &host_mapped_dev_ptr, sizeof(host_mapped_dev_ptr)));
  CFB_CUDA(cudaMemcpyToSymbol(g_wait_limit, &wait_limit_cycles, sizeof(wait_limit_cycles)));
  CFB_CUDA(cudaMemcpyToSymbol(g_abort, &zero, sizeof(zero)));
  return 0;
}
static std::atomic<int> g_inject_fault{0};
static std::atomic<long long*> g_stamps{nullptr};      // diagnostics: device buffer of >= 32 int64 (cfb_debug_set_stamps)
int tc_set_stamps(long long* dev_ptr) { g_stamps.store(dev_ptr); return 0; }
int tc_inject_fault(int kind) { g_inject_fault.store(kind); return 0; }
int tc_clear_abort() {
  const unsigned zero = 0;
  CFB_CUDA(cudaMemcpyToSymbol(g_abort, &zero, sizeof(zero)));
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

static int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int spatial_stride = 1, bool f32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  CFB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  // traversal stride 2 along W and H turns the box into the stride-2 sampling pattern of Downsample
  cuuint32_t estr[5] = {1, (cuuint32_t)spatial_stride, (cuuint32_t)spatial_stride, 1, 1};
  const CUresult rc = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                         reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                         reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CFB_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")");
  return 0;
}

static inline int tile_bw(int Wo) { return Wo < 128 ? Wo : 128; }

// k-blocks (64 K-elements) per TMEM partial sum; env CFB_TC_CHUNK overrides for experiments (1..64)
static int tc_chunk_kblocks() {
  static int v = [] {
    const char* e = getenv("CFB_TC_CHUNK");
    int c = e ? atoi(e) : 8;
    return c < 1 ? 1 : (c > 4096 ? 4096 : c);
  }();
  return v;
}

// CTA-pair engine on unless CFB_TC_PAIR=0 (experiments / A-B timing)
static bool pair_enabled() {
  static const bool v = [] { const char* e = getenv("CFB_TC_PAIR"); return !(e && atoi(e) == 0); }();
  return v;
}
// halo engine (one patch fetch per 64-channel block, taps via shifted descriptors): 3x3 stride-1 convs and the 2x2 parity
// convs of Upsample, on 8x16-pixel tiles.  Measured on B200 (profiles/round1_*): with single-CTA MMAs it is SLOWER than the
// per-tap engine (the row-shifted A views make the A fetch 10-25 % slower and the single-CTA MMA floor is the limit); in
// CTA-pair mode the MMA has headroom and the step is power-capped, so the halved L2->SM traffic wins (+13 % on the dominant
// conv, +5 % on the step).  It is therefore used exactly when the pair engine is (CFB_TC_HALO=0 / CFB_TC_PAIR=0 switch it off).
static bool halo_enabled() {
  static bool v = [] { const char* e = getenv("CFB_TC_HALO"); return !(e && atoi(e) == 0); }();
  return v;
}
// 64-wide n-tiles for few-tile launches of wider layers (CFB_TC_SMALLN=0 keeps 128)
static bool small_n_enabled() {
  static const bool v = [] { const char* e = getenv("CFB_TC_SMALLN"); return !(e && atoi(e) == 0); }();
  return v;
}
struct TcGeom { int BW, BH; bool halo; };
static TcGeom tc_geometry(const ConvArgs& a) {
  TcGeom g;
  const int Wt = a.mode == CONV_UP ? a.W : a.Wo, Ht = a.mode == CONV_UP ? a.H : a.Ho;   // grid the tiles live on
  g.halo = halo_enabled() && pair_enabled() && (a.ksize == 3 || (a.ksize == 1 && a.halo1x1 && a.mode == CONV_SAME)) &&
           (a.mode == CONV_SAME || a.mode == CONV_UP) &&
           ((Wt % 8 == 0 && Ht % 16 == 0) || a.gen);      // gen: ragged tiles, stores are bounds-checked
  if (g.halo) { g.BW = 8; g.BH = 16; }
  else { g.BW = tile_bw(a.mode == CONV_UP ? a.W : a.Wo); g.BH = 128 / g.BW; }   // Upsample: tiles live on the low-res grid
  return g;
}

int tc_tiles_per_image(const ConvArgs& a) {
  const TcGeom g = tc_geometry(a);
  if (a.mode == CONV_UP) return 4 * (a.W / g.BW) * (a.H / g.BH);
  return (a.Wo / g.BW) * (a.Ho / g.BH);
}

bool tc_supported(const ConvArgs& a) {
  if (a.Cin % 64 != 0 || a.Cout % 64 != 0) return false;
  if (!(a.ksize == 1 || a.ksize == 3)) return false;
  if (a.mode == CONV_DOWN && a.ksize != 3) return false;
  if (a.Wo < 1 || a.Ho < 1) return false;
  const TcGeom g = tc_geometry(a);
  if (a.gen) return g.halo && a.ksize == 3;
  const int BW = g.BW, BH = g.BH;
  const int Wt = a.mode == CONV_UP ? a.W : a.Wo, Ht = a.mode == CONV_UP ? a.H : a.Ho;   // grid the tiles live on
  if (a.mode == CONV_UP && a.ksize != 3) return false;
  if (128 % BW != 0 || Wt % BW != 0) return false;
  if (Ht % BH != 0) return false;
  if ((int64_t)a.N * (a.Ho / BH + 1) * (a.Wo / BW + 1) * (a.Cout / 64) > 0x7fffffffLL) return false;
  return true;
}

// fused operand transform: available for 3x3 stride-1 convs on the halo + pair engine.  Round 1 read fp16 hi/lo planes and
// only paid on the 128-wide layers at >= 64x64; the round-2 transform reads the fp32 activation itself (no planes written by
// the producer, no prep pass), needs ~3x fewer instructions per element and has 8 transform warps + 3 patch slots on the
// 64-wide layers, so it is used wherever the engine exists.  CFB_TC_XFORM=0 disables it, =3 restores the round-1 rule.
bool tc_can_xform(const ConvArgs& a) {
  static const int mode = [] { const char* e = getenv("CFB_TC_XFORM"); return e ? atoi(e) : 1; }();
  if (a.gen) return tc_supported(a);      // generalised variant: tile count is padded to an even number, CONV_UP included
  if (mode == 0 || !tc_supported(a) || a.mode != CONV_SAME || !(a.ksize == 3 || (a.ksize == 1 && a.halo1x1 && a.Cout % 128 == 0))) return false;
  if (mode == 3 && !(a.Cout % 128 == 0 && a.Cin >= 128 && (int64_t)a.Ho * a.Wo >= 4096)) return false;
  const TcGeom g = tc_geometry(a);
  if (!g.halo) return false;
  const int64_t m_tiles = (int64_t)a.N * (a.Wo / g.BW) * (a.Ho / g.BH);
  return m_tiles % 2 == 0;
}

size_t tc_scratch_bytes(const ConvArgs& a) {
  if (!tc_supported(a)) return 0;
  const int Hp = a.mode == CONV_SAME ? a.Ho : a.H, Wp = a.mode == CONV_SAME ? a.Wo : a.W;   // operand plane = input resolution
  const size_t plane = ((size_t)a.N * Hp * Wp * a.Cin * 2 + 1023) / 1024 * 1024;
  return 2 * plane;
}

// pair mode needs both CTAs of a pair on m-tiles that share the weight slice: an even number of m-tiles
// (Upsample: of low-resolution tiles, i.e. m_tiles % 8 == 0)
static bool pair_ok(const TcParams& p) {
  if (!pair_enabled()) return false;
  return p.up4 ? (p.m_tiles % 8 == 0) : (p.m_tiles % 2 == 0);
}

struct TcMaps { CUtensorMap a_hi, a_lo, b_hi, b_lo, b_half; };

template <int BN, int CPG, bool HALO, bool PAIR, bool XF = false, bool GEN = false, bool K1 = false>
static int launch_tc2(const TcMaps& m, const TcParams& p, int sm_count, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  constexpr int SMEM = XF ? Cfg::X_SMEM_BYTES
                          : (HALO ? (PAIR ? Cfg::HP_SMEM_BYTES : Cfg::H_SMEM_BYTES) : (PAIR ? Cfg::P_SMEM_BYTES : Cfg::SMEM_BYTES));
  constexpr int THREADS = XF ? TcCfg<BN>::XF_THREADS : TC_THREADS;
  static_assert(SMEM <= 232448, "shared memory budget");
  // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of the function: remember it per device
  // (one process may drive several GPUs from several threads)
  static std::atomic<uint64_t> attr_done{0};
  int dev = 0;
  CFB_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    CFB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, CPG, HALO, PAIR, XF, GEN, K1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  if constexpr (PAIR) {
    const int pairs = (p.m_tiles / 2) * p.n_tiles;
    const int max_pairs = sm_count / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * (pairs < max_pairs ? pairs : max_pairs)));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    CFB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, CPG, HALO, PAIR, XF, GEN, K1>, m.a_hi, m.a_lo, m.b_hi, m.b_lo, m.b_half, p));
    count_launch();
  } else {
    const int total = p.m_tiles * p.n_tiles;
    const int grid = total < sm_count ? total : sm_count;
    CFB_LAUNCH_PDL((conv_tc_kernel<BN, CPG, HALO, PAIR, XF, GEN, K1>), dim3((unsigned)grid), dim3(THREADS), (size_t)SMEM, st, m.a_hi, m.a_lo,
                   m.b_hi, m.b_lo, m.b_half, p);
  }
  return 0;
}
template <int BN, int CPG>
static int launch_tc(const TcMaps& m, const TcParams& p, int sm_count, cudaStream_t st, bool gen = false) {
  if constexpr (CPG == 0) {
    if (gen) {
      CFB_REQUIRE(p.xform && p.PW == 10 && p.PH == 18 && p.m_tiles % 2 == 0, "conv_tc: generalised variant needs the halo + pair + transform engine");
      return launch_tc2<BN, 0, true, true, true, true>(m, p, sm_count, st);
    }
  }
  if (p.xform) {         // fused operand transform: conv_tc() only asks for it when tc_can_xform() holds
    CFB_REQUIRE(((p.PW == 10 && p.PH == 18) || (p.PW == 8 && p.PH == 16 && p.taps == 1)) && pair_ok(p),
                "conv_tc: fused operand transform needs the halo + pair engine");
    if (p.taps == 1) {      // 1x1 conv with a GroupNorm-affine input (AttnBlock q,k,v): patch = tile
      if constexpr (BN == 128 && CPG == 0) return launch_tc2<128, 0, true, true, true, false, true>(m, p, sm_count, st);
      else { CFB_REQUIRE(false, "conv_tc: the 1x1 fused transform is built for 128-wide tiles without statistics"); }
    }
    return launch_tc2<BN, CPG, true, true, true>(m, p, sm_count, st);
  }
  if (p.PW > 0) return pair_ok(p) ? launch_tc2<BN, CPG, true, true>(m, p, sm_count, st) : launch_tc2<BN, CPG, true, false>(m, p, sm_count, st);
  if (pair_ok(p)) return launch_tc2<BN, CPG, false, true>(m, p, sm_count, st);
  return launch_tc2<BN, CPG, false, false>(m, p, sm_count, st);
}

// GroupNorm partial sums can be emitted for (BN=64, Cout=64) and (BN=128, Cout in {128,256,512})
bool tc_can_emit_stats(const ConvArgs& a) {
  if (!tc_supported(a)) return false;
  if (a.Cout % 128 == 0) return a.Cout == 128 || a.Cout == 256 || a.Cout == 512;
  return a.Cout == 64;
}

int conv_tc(const ConvArgs& a, void* scratch, int sm_count, cudaStream_t st) {
  CFB_REQUIRE(tc_supported(a), "conv_tc: unsupported shape");
  CFB_REQUIRE(a.wgt_hi && a.wgt_lo && a.wscale_inv, "conv_tc: split weights missing");
  const int64_t M = (int64_t)a.N * a.Ho * a.Wo;
  if (M == 0) return 0;
  // ---- operand planes (fp16 hi/lo NHWC; at the output resolution, or the input resolution for Downsample)
  const int Hp = a.mode == CONV_SAME ? a.Ho : a.H, Wp = a.mode == CONV_SAME ? a.Wo : a.W;
  const int64_t Mp = (int64_t)a.N * Hp * Wp;
  const size_t plane = ((size_t)Mp * a.Cin * 2 + 1023) / 1024 * 1024;
  __half* hi = (__half*)scratch;
  __half* lo = (__half*)((char*)scratch + plane);
  if (!a.skip_prep) {
    const int C8 = a.Cin / 8;
    CFB_REQUIRE(C8 <= 256 && 256 % C8 == 0, "conv_tc: Cin must be 64 * 2^k (<= 2048)");
    const int64_t img_px = (int64_t)Hp * Wp;
    // pixels per thread: up to 32 (amortises the per-block affine loads) but never so many that the grid drops
    // below ~16 blocks per SM -- the small 16x16 / 32x32 layers are latency-bound otherwise
    const int pstep = 256 / C8;
    int iters = 32;
    while (iters > 1 && Mp / ((int64_t)pstep * iters) < 148 * 16) iters >>= 1;
    int64_t PB = (int64_t)pstep * iters;
    while (PB > 1 && img_px % PB != 0) PB >>= 1;
    CFB_REQUIRE(PB >= 1 && img_px % PB == 0 && Mp % PB == 0, "conv_tc: image size not supported by the operand prep kernel");
    tc_prep_kernel<<<(unsigned)(Mp / PB), 256, 0, st>>>(a.in, a.in_scale, a.in_shift, a.in_act, 0, a.N, a.H,
                                                        a.W, a.Cin, (int)PB, hi, lo);
    CFB_LAUNCH_CHECK();
  }
  // ---- tensor maps
  const TcGeom geo = tc_geometry(a);
  const int BW = geo.BW, BH = geo.BH;
  const int PW = geo.halo ? BW + a.ksize - 1 : 0, PH = geo.halo ? BH + a.ksize - 1 : 0;
  int BN = (a.Cout % 128 == 0) ? 128 : 64;
  // Few-tile launches (a single face; the 16x16 .. 64x64 layers of a small batch): with 128-wide n-tiles a handful of CTA pairs
  // carry the whole K loop while most SMs idle (512 -> 512 @16x16, one face: 4 pairs, 30 us of MMAs each).  64-wide n-tiles
  // double the number of pairs.  Every output element keeps its accumulation order and every GroupNorm partial its summation
  // tree, so the result is bit-identical whichever width runs (the batch-invariance tests cover both).
  if (BN == 128 && a.xform && !a.gen && a.ksize == 3 && a.mode == CONV_SAME && geo.halo && small_n_enabled()) {
    const int64_t pairs128 = ((int64_t)a.N * (a.Wo / BW) * (a.Ho / BH) / 2) * (a.Cout / 128);
    if (pairs128 * 2 <= sm_count / 2) BN = 64;
  }
  TcMaps mp;
  CUtensorMap &mA_hi = mp.a_hi, &mA_lo = mp.a_lo, &mB_hi = mp.b_hi, &mB_lo = mp.b_lo;
  if (a.xform) {
    // fused operand transform: the A operand is read straight from the fp32 NHWC activation(s); boxes of 32 channels
    // (128 B rows) x the halo patch.  Source 1 is the second half of a channel concatenation (or source 0 again).
    CFB_REQUIRE(a.in != nullptr && geo.halo, "conv_tc: fused operand transform needs the fp32 input and the halo engine");
    // gen: `in` points into a wider NHWC buffer of in_pitch channels; the 64-aligned window [0, Cin) is read (channels beyond
    // the buffer are zero-filled by the TMA unit, channels beyond the real Cin meet zero weights)
    const int C0 = a.in2 ? a.Cin1 : (a.gen && a.in_pitch ? a.in_pitch : a.Cin), C1 = a.in2 ? a.Cin - a.Cin1 : C0;
    CFB_REQUIRE(!(a.gen && a.in2), "conv_tc: the generalised variant reads one source");
    CFB_REQUIRE(a.gen || (C0 % 64 == 0 && C1 % 64 == 0 && C0 > 0 && C1 > 0), "conv_tc: concatenated sources must be multiples of 64 channels");
    CFB_REQUIRE(C0 % 4 == 0 && C0 > 0, "conv_tc: source channel pitch must be a multiple of 4");
    const uint32_t box[4] = {32, (uint32_t)PW, (uint32_t)PH, 1};
    {
      const uint64_t dims[4] = {(uint64_t)C0, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)a.N};
      const uint64_t str[3] = {(uint64_t)C0 * 4, (uint64_t)Wp * C0 * 4, (uint64_t)Hp * Wp * C0 * 4};
      CFB_CHECK(make_map(&mA_hi, a.in, 4, dims, str, box, 1, true));
    }
    {
      const uint64_t dims[4] = {(uint64_t)C1, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)a.N};
      const uint64_t str[3] = {(uint64_t)C1 * 4, (uint64_t)Wp * C1 * 4, (uint64_t)Hp * Wp * C1 * 4};
      CFB_CHECK(make_map(&mA_lo, a.in2 ? a.in2 : a.in, 4, dims, str, box, 1, true));
    }
  } else {
    const int sp = a.mode == CONV_DOWN ? 2 : 1;
    const uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)a.N};
    const uint64_t str[3] = {(uint64_t)a.Cin * 2, (uint64_t)Wp * a.Cin * 2, (uint64_t)Hp * Wp * a.Cin * 2};
    // per-tap engine: ceil(box/stride) = BW x BH pixels land; halo engine: the whole (BW+k-1) x (BH+k-1) patch
    const uint32_t box[4] = {64, (uint32_t)(geo.halo ? PW : BW * sp), (uint32_t)(geo.halo ? PH : BH * sp), 1};
    CFB_CHECK(make_map(&mA_hi, hi, 4, dims, str, box, sp));
    CFB_CHECK(make_map(&mA_lo, lo, 4, dims, str, box, sp));
  }
  {
    const int taps = a.mode == CONV_UP ? 16 : a.ksize * a.ksize;
    const uint64_t dims[3] = {(uint64_t)a.Cin, (uint64_t)a.Cout, (uint64_t)taps};
    const uint64_t str[2] = {(uint64_t)a.Cin * 2, (uint64_t)a.Cout * a.Cin * 2};
    const uint32_t box[3] = {64, (uint32_t)BN, 1};
    CFB_CHECK(make_map(&mB_hi, a.wgt_hi, 3, dims, str, box));
    CFB_CHECK(make_map(&mB_lo, a.wgt_lo, 3, dims, str, box));
    const uint32_t hbox[3] = {64, (uint32_t)(BN / 2), 1};      // CTA-pair engine: each CTA stages half of B_hi again
    CFB_CHECK(make_map(&mp.b_half, a.wgt_hi, 3, dims, str, hbox));
  }
  TcParams p;
  p.N = a.N; p.Ho = a.Ho; p.Wo = a.Wo; p.Cout = a.Cout;
  p.taps = a.ksize * a.ksize; p.pad = a.mode == CONV_DOWN ? 0 : a.ksize / 2; p.stride = a.mode == CONV_DOWN ? 2 : 1;
  p.up4 = a.mode == CONV_UP ? 1 : 0;
  p.a_c0 = 0; p.b_c0 = 0; p.b_batched = 0;
  p.heads = 1; p.a_c_head = 0; p.b_c_head = 0; p.a_img_per_head = 0; p.b_r_head = 0; p.out_per_head = 1; p.o_c_head = 0;
  if (p.up4) { p.taps = 4; p.pad = 1; }
  p.chunk = tc_chunk_kblocks();
  p.PW = PW; p.PH = PH;
  p.BW = BW; p.BH = BH;
  p.tiles_x = ((p.up4 ? a.W : a.Wo) + BW - 1) / BW; p.tiles_y = ((p.up4 ? a.H : a.Ho) + BH - 1) / BH;   // exact unless gen (ragged)
  {
    int64_t lowres_tiles = (int64_t)a.N * p.tiles_x * p.tiles_y;
    if (a.gen) lowres_tiles += lowres_tiles & 1;      // CTA pairs: an even tile count (the dummy tile stores nothing)
    p.m_tiles = (int)(lowres_tiles * (p.up4 ? 4 : 1));
  }
  p.n_tiles = a.Cout / BN; p.kblocks = a.Cin / 64;
  p.Hin = Hp; p.Win = Wp; p.pad_mode = a.pad_mode; p.sub = a.subsample ? 1 : 0;
  p.out_pitch = a.out_pitch ? a.out_pitch : a.Cout; p.out_c0 = a.out_c0; p.cout_valid = a.cout_valid ? a.cout_valid : a.Cout;
  p.res_pitch = a.res_pitch ? a.res_pitch : p.out_pitch; p.residual2 = a.residual2;
  p.res2_pitch = a.res2_pitch ? a.res2_pitch : p.out_pitch; p.post_scale = a.post_scale;
  if (a.gen) {
    CFB_REQUIRE(a.xform && !a.gn_part && !a.out_planes && !a.sft_dec, "conv_tc: generalised variant = fused transform, fp32 output only");
    CFB_REQUIRE(p.out_pitch % 4 == 0 && p.out_c0 % 4 == 0 && p.cout_valid % 4 == 0 && p.res_pitch % 4 == 0 && p.res2_pitch % 4 == 0,
                "conv_tc: channel pitches / offsets must be multiples of 4");
    CFB_REQUIRE(!a.subsample || (a.mode == CONV_SAME && a.Ho % 2 == 0 && a.Wo % 2 == 0), "conv_tc: subsampling needs even sizes");
    CFB_REQUIRE(a.out_act == OUT_NONE || a.out_act == OUT_LRELU, "conv_tc: generalised variant has bias / residual / LeakyReLU epilogues");
  }
  if (p.taps * p.kblocks <= 12) p.chunk = p.taps * p.kblocks;   // short K (Cin = 64): one partial sum, no 8+1 split
  p.in_scale = nullptr; p.in_shift = nullptr; p.in_act = IN_NONE; p.xform = a.xform ? 1 : 0;
  p.fault = g_inject_fault.exchange(0);
#if CFB_TC_STAMPS
  p.dbg = g_stamps.load();
#endif
  p.vq_e2 = a.vq_e2; p.vq_z2 = a.vq_z2; p.vq_cand = a.vq_cand; p.vq_dpart = a.vq_dpart;
  CFB_REQUIRE(!a.vq_cand || (a.vq_e2 && a.vq_z2 && a.vq_dpart && a.ksize == 1 && !a.gen && !a.xform), "conv_tc: VQ argmin epilogue needs e2, z2 and the 1x1 engine");
  p.a_split = a.in2 ? a.Cin1 / 64 : a.Cin / 64;
  if (a.xform) {
    CFB_REQUIRE(a.skip_prep && tc_can_xform(a), "conv_tc: fused operand transform not available for this conv");
    CFB_REQUIRE((a.in_scale != nullptr) == (a.in_shift != nullptr) && (a.in_scale || a.in_act == IN_NONE),
                "conv_tc: fused operand transform takes scale and shift together");
    p.in_scale = a.in_scale; p.in_shift = a.in_shift; p.in_act = a.in_act;
  } else {
    CFB_REQUIRE(a.in2 == nullptr, "conv_tc: a two-source input needs the fused operand transform");
  }
  p.bias = a.bias; p.residual = a.residual; p.out_act = a.out_act;
  p.sft_dec = a.sft_dec; p.sft_scale = a.sft_scale; p.sft_w = a.sft_w; p.wscale_inv = a.wscale_inv; p.out = a.out;
  p.gn_part = a.gn_part; p.gn_cpg = a.Cout / 32;
  p.pl_hi = (__half*)a.out_planes;
  p.pl_lo = a.out_planes ? (__half*)((char*)a.out_planes + (((size_t)a.N * a.Ho * a.Wo * a.Cout * 2 + 1023) / 1024 * 1024)) : nullptr;
  const int cpg = a.gn_part ? a.Cout / 32 : 0;
  CFB_REQUIRE(!a.gn_part || tc_can_emit_stats(a), "conv_tc: GroupNorm partials are not available for this Cout");
  if (a.gen) return BN == 128 ? launch_tc<128, 0>(mp, p, sm_count, st, true) : launch_tc<64, 0>(mp, p, sm_count, st, true);
  if (BN == 128) {
    switch (cpg) {
      case 0: return launch_tc<128, 0>(mp, p, sm_count, st);
      case 4: return launch_tc<128, 4>(mp, p, sm_count, st);
      case 8: return launch_tc<128, 8>(mp, p, sm_count, st);
      case 16: return launch_tc<128, 16>(mp, p, sm_count, st);
    }
  } else {
    if (cpg == 0) return launch_tc<64, 0>(mp, p, sm_count, st);
    if (cpg == 2) return launch_tc<64, 2>(mp, p, sm_count, st);
    // 64-wide tiles of a wider layer (few-tile launches, see above): fused-transform engine only
    if (p.xform && p.taps == 9 && p.PW == 10 && p.PH == 18 && pair_ok(p)) {
      if (cpg == 4) return launch_tc2<64, 4, true, true, true>(mp, p, sm_count, st);
      if (cpg == 8) return launch_tc2<64, 8, true, true, true>(mp, p, sm_count, st);
      if (cpg == 16) return launch_tc2<64, 16, true, true, true>(mp, p, sm_count, st);
    }
  }
  CFB_REQUIRE(false, "conv_tc: no kernel variant for this configuration");
  return 1;
}


// ------------------------------------------------------------------------------------------------------
// Batched GEMM on the same engine (attention cores): per image n
//     out[n][t][j] = scale * sum_k A[n][t][a_c0 + k] * B[n][j][b_c0 + k],   t in 0..255 (16x16 tokens), j in 0..Cout-1
// A and B are fp16 hi/lo operand planes (token-major, channel pitch a_pitch / b_pitch); B is addressed per image
// through the third TMA coordinate.  Used for  scores = q k^T * C^-1/2  and  out = P v  of AttnBlock
// (/root/reference/basicsr/archs/vqgan_arch.py:209-222).
// ------------------------------------------------------------------------------------------------------
int bmm_tc(const BmmArgs& g, int sm_count, cudaStream_t st) {
  const int BN = g.Cout % 128 == 0 ? 128 : 64;
  CFB_REQUIRE(g.K % 64 == 0 && g.Cout % 64 == 0 && g.N >= 0 && g.heads >= 1, "bmm_tc: K and Cout must be multiples of 64");
  CFB_REQUIRE(g.a_c0 % 64 == 0 && g.b_c0 % 64 == 0 && g.a_pitch % 8 == 0 && g.b_pitch % 8 == 0 && g.a_c_head % 64 == 0 &&
                  g.b_c_head % 64 == 0 && g.b_r_head % BN == 0 && g.o_c_head % 4 == 0,
              "bmm_tc: unaligned operand slice");
  if (g.N == 0) return 0;
  const int a_imgs = g.a_img_per_head ? g.N * g.heads : g.N;
  const size_t a_plane = ((size_t)a_imgs * 256 * g.a_pitch * 2 + 1023) / 1024 * 1024;
  const size_t b_plane = ((size_t)g.N * g.b_rows * g.b_pitch * 2 + 1023) / 1024 * 1024;
  TcMaps mp;
  CUtensorMap &mA_hi = mp.a_hi, &mA_lo = mp.a_lo, &mB_hi = mp.b_hi, &mB_lo = mp.b_lo;
  {
    const uint64_t dims[4] = {(uint64_t)g.a_pitch, 16, 16, (uint64_t)a_imgs};
    const uint64_t str[3] = {(uint64_t)g.a_pitch * 2, (uint64_t)16 * g.a_pitch * 2, (uint64_t)256 * g.a_pitch * 2};
    const uint32_t box[4] = {64, 16, 8, 1};
    CFB_CHECK(make_map(&mA_hi, g.a_planes, 4, dims, str, box));
    CFB_CHECK(make_map(&mA_lo, (const char*)g.a_planes + a_plane, 4, dims, str, box));
  }
  {
    const uint64_t dims[3] = {(uint64_t)g.b_pitch, (uint64_t)g.b_rows, (uint64_t)g.N};
    const uint64_t str[2] = {(uint64_t)g.b_pitch * 2, (uint64_t)g.b_rows * g.b_pitch * 2};
    const uint32_t box[3] = {64, (uint32_t)BN, 1};
    CFB_CHECK(make_map(&mB_hi, g.b_planes, 3, dims, str, box));
    CFB_CHECK(make_map(&mB_lo, (const char*)g.b_planes + b_plane, 3, dims, str, box));
    const uint32_t hbox[3] = {64, (uint32_t)(BN / 2), 1};
    CFB_CHECK(make_map(&mp.b_half, g.b_planes, 3, dims, str, hbox));
  }
  const int out_pitch = g.out_per_head ? g.Cout : g.heads * g.o_c_head;      // channels per token row of `out`
  const int out_imgs = g.out_per_head ? g.N * g.heads : g.N;
  CFB_REQUIRE(g.heads == 1 || g.out_per_head || g.o_c_head == g.Cout, "bmm_tc: a head's column slice must equal its Cout");
  TcParams p;
  p.N = g.N * g.heads; p.Ho = 16; p.Wo = 16; p.Cout = out_pitch;
  p.taps = 1; p.pad = 0; p.stride = 1; p.up4 = 0;
  p.a_c0 = g.a_c0; p.b_c0 = g.b_c0; p.b_batched = 1;
  p.heads = g.heads; p.a_c_head = g.a_c_head; p.b_c_head = g.b_c_head; p.a_img_per_head = g.a_img_per_head ? 1 : 0;
  p.b_r_head = g.b_r_head; p.out_per_head = g.out_per_head ? 1 : 0; p.o_c_head = g.o_c_head;
  p.chunk = tc_chunk_kblocks();
  p.PW = 0; p.PH = 0;
  p.BW = 16; p.BH = 8; p.tiles_x = 1; p.tiles_y = 2;
  p.m_tiles = g.N * g.heads * 2; p.n_tiles = g.Cout / BN; p.kblocks = g.K / 64;
  if (p.kblocks <= 12) p.chunk = p.kblocks;
  p.in_scale = nullptr; p.in_shift = nullptr; p.in_act = IN_NONE; p.xform = 0; p.a_split = 0; p.fault = 0;
#if CFB_TC_STAMPS
  p.dbg = g_stamps.load();
#endif
  p.vq_e2 = nullptr; p.vq_z2 = nullptr; p.vq_cand = nullptr; p.vq_dpart = nullptr;
  p.Hin = 16; p.Win = 16; p.pad_mode = 0; p.sub = 0; p.out_pitch = out_pitch; p.out_c0 = 0; p.cout_valid = out_pitch; p.res_pitch = out_pitch;
  p.residual2 = nullptr; p.res2_pitch = out_pitch; p.post_scale = 1.f;
  p.bias = nullptr; p.residual = nullptr; p.out_act = OUT_NONE; p.sft_dec = nullptr; p.sft_scale = nullptr; p.sft_w = 0.f;
  p.wscale_inv = g.scale_dev; p.out = g.out;
  p.gn_part = nullptr; p.gn_cpg = 0;
  p.pl_hi = (__half*)g.out_planes;
  p.pl_lo = g.out_planes ? (__half*)((char*)g.out_planes + (((size_t)out_imgs * 256 * out_pitch * 2 + 1023) / 1024 * 1024)) : nullptr;
  return BN == 128 ? launch_tc<128, 0>(mp, p, sm_count, st) : launch_tc<64, 0>(mp, p, sm_count, st);
}

// softmax over rows of 256 fp32 scores -> fp16 hi/lo operand planes of the probabilities (one warp per row)
__global__ void __launch_bounds__(256) softmax256_planes_kernel(const float* __restrict__ s, __half* __restrict__ hi,
                                                                __half* __restrict__ lo, int64_t rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int l = threadIdx.x & 31;
  if (row >= rows) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(s + row * 256 + l * 8));
  const float4 b = __ldg(reinterpret_cast<const float4*>(s + row * 256 + l * 8 + 4));
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float mx = v[0];
#pragma unroll
  for (int j = 1; j < 8; ++j) mx = fmaxf(mx, v[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = expf(v[j] - mx); sum += v[j]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  __align__(16) __half hh[8];
  __align__(16) __half ll[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float pv = v[j] * inv;
    hh[j] = __float2half_rn(pv);
    ll[j] = __float2half_rn(pv - __half2float(hh[j]));
  }
  *reinterpret_cast<uint4*>(hi + row * 256 + l * 8) = *reinterpret_cast<const uint4*>(hh);
  *reinterpret_cast<uint4*>(lo + row * 256 + l * 8) = *reinterpret_cast<const uint4*>(ll);
}
int softmax256_planes(const float* scores, void* planes, int64_t rows, cudaStream_t st) {
  if (rows == 0) return 0;
  const size_t plane = ((size_t)rows * 256 * 2 + 1023) / 1024 * 1024;
  CFB_LAUNCH_PDL(softmax256_planes_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, scores, (__half*)planes,
                 (__half*)((char*)planes + plane), rows);
  return 0;
}

// V^T operand planes: in planes [N][256 tokens][pitch] (channels c0..c0+C) -> out planes [N][C][256], hi and lo
__global__ void transpose_planes_kernel(const __half* __restrict__ in, __half* __restrict__ out, int pitch, int c0, int C) {
  __shared__ __half tile[32][34];
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.z, t0 = blockIdx.y * 32, cb = blockIdx.x * 32;
  const __half* ib = in + (int64_t)n * 256 * pitch;
  __half* ob = out + (int64_t)n * C * 256;
  for (int i = threadIdx.y; i < 32; i += 8) tile[i][threadIdx.x] = ib[(int64_t)(t0 + i) * pitch + c0 + cb + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) ob[(int64_t)(cb + i) * 256 + t0 + threadIdx.x] = tile[threadIdx.x][i];
}
int transpose_planes(const void* in_planes, int N, int pitch, int c0, int C, void* out_planes, cudaStream_t st) {
  CFB_REQUIRE(C % 32 == 0, "transpose_planes: C must be a multiple of 32");
  if (N == 0) return 0;
  const size_t ip = ((size_t)N * 256 * pitch * 2 + 1023) / 1024 * 1024;
  const size_t op = ((size_t)N * C * 256 * 2 + 1023) / 1024 * 1024;
  dim3 grid(C / 32, 8, N);
  CFB_REQUIRE(N <= 65535, "transpose_planes: batch too large");
  for (int h = 0; h < 2; ++h) {
    CFB_LAUNCH_PDL(transpose_planes_kernel, grid, dim3(32, 8), 0, st, (const __half*)((const char*)in_planes + h * ip),
                   (__half*)((char*)out_planes + h * op), pitch, c0, C);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// VectorQuantizer.forward as ONE kernel (BASELINE configs[2]; /root/reference/basicsr/archs/vqgan_arch.py:33-70)
//   d = |z|^2 + |e|^2 - 2 z.E^T ; argmin ; z_q = z + (E[idx] - z) ; loss ; perplexity ; mean_distance
// on the caller's NCHW tensors.  One cluster of two CTAs owns 256 consecutive tokens of one image (128 per CTA) and ALL codes:
//   phase 0  every thread reads its (token, 8-channel) items of z straight from NCHW (coalesced along the tokens), splits them
//            into fp16 hi / lo and writes them into shared memory in the K-major 128B-swizzled layout the MMA descriptors read
//            (the whole 128 x 256 A tile of the CTA stays resident: 128 KB), accumulates |z|^2;
//   phase 1  warp 0 streams the prepared codebook planes ([B_hi ; B_lo] halves per CTA + its half of B_hi, as in the conv engine)
//            through a 3-stage TMA ring, warp 1 of the leader issues tcgen05.mma.cta_group::2 (M = 256) into a double-buffered
//            TMEM slot per 128-code chunk, warps 2..5 read the finished slot and keep (first minimum, index) per token row;
//   phase 2  all threads gather E[idx], write the straight-through z_q in NCHW and the squared-error partial sums; the last CTA
//            of the grid (ticket) turns the partial sums and the code histogram into the three statistics and clears them.
// The [tokens, codes] distance matrix never leaves TMEM; no other kernel, copy or transpose runs.
// ------------------------------------------------------------------------------------------------------
struct VqParams {
  const float* z; const float* codebook; const float* e2; const float* wscale_inv;
  float* zq; int64_t* idx; float* stats;
  double* part;          // [ctas][2]: squared error, sum of all distances
  unsigned* hist;        // [K]: zero on entry, cleared again by the last CTA
  unsigned* ticket;      // zero on entry
  int N, D, HW, K, kblocks, nchunks;
  float beta;
  long long* dbg;        // optional [ctas][6] clock64 stamps of the phase boundaries (tools/gpu_r2_vq.sh)
};
constexpr int VQ_THREADS = 512;      // 16 warps: 0 TMA, 1 MMA, 2..5 epilogue; all 16 move z / z_q in phases 0 and 2
constexpr int VQ_BX = 128 * 128, VQ_BY = 64 * 128, VQ_STAGE = VQ_BX + VQ_BY, VQ_STAGES = 3;
constexpr int VQ_A_KB = 128 * 128;      // one 64-channel k-block of one plane: 128 token rows x 128 B

__global__ void __launch_bounds__(VQ_THREADS, 1)
vq_fused_kernel(const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ CUtensorMap tmB_half, const VqParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_hi = smem;                                   // [kblocks][128][128 B]
  uint8_t* a_lo = smem + 4 * VQ_A_KB;
  uint8_t* ring = smem + 8 * VQ_A_KB;                     // VQ_STAGES x (X | Y)
  float* e2s = reinterpret_cast<float*>(ring + VQ_STAGES * VQ_STAGE);      // [K <= 1024]
  float* z2p = e2s + 1024;                                // [4 k-blocks][128 tokens]
  int* bidx = reinterpret_cast<int*>(z2p + 512);          // [128]
  double* red = reinterpret_cast<double*>(bidx + 128);    // [8][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 16);
  uint64_t* full = bars;
  uint64_t* empty = bars + VQ_STAGES;
  uint64_t* cfull = bars + 2 * VQ_STAGES;
  uint64_t* cempty = cfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cempty + 2);
  int* s_flags = reinterpret_cast<int*>(tmem_slot + 1);   // [0] abort seen, [1] last CTA

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cl = blockIdx.x >> 1;
  const int per_img = p.HW / 256;
  const int n = cl / per_img;
  const int hw0 = (cl - n * per_img) * 256 + (int)rank * 128;
  bool aborted = false;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 6 + 0] = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < VQ_STAGES; ++s) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(cfull + a), 1); mbar_init(smem_u32(cempty + a), 8); }
    s_flags[0] = 0; s_flags[1] = 0;
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_half) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // ---- phase 0: z (NCHW) -> fp16 hi / lo operand tile in shared memory, |z|^2, |e|^2 table
  // One warp item = 16 tokens x one 64-channel k-block: lane (a = lane>>3, b = lane&7) reads 8 channels (b*8..) of 4 consecutive
  // tokens (a) with 16-byte loads along the token axis (64 B segments: every sector fully used) and writes, per token, the
  // 16-byte chunk b of that token's row -- 8 lanes cover a 128 B row, 4 rows per instruction: conflict-free STS.128.
  {
    const int la = lane >> 3, lb = lane & 7;
    for (int k = threadIdx.x; k < p.K; k += VQ_THREADS) e2s[k] = __ldg(p.e2 + k);
    const int items = 8 * p.kblocks;                      // 8 groups of 16 tokens x kblocks
    const float* zn = p.z + (int64_t)n * p.D * p.HW + hw0;
    for (int it0 = warp; it0 < items; it0 += 32) {        // two items in flight per warp (all of them at D = 256)
      float4 v[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int item = it0 + 16 * u;
        if (item < items) {
          const int kb = item % p.kblocks, t4 = ((item / p.kblocks) * 4 + la) * 4;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[u][j] = __ldg(reinterpret_cast<const float4*>(zn + (int64_t)(kb * 64 + lb * 8 + j) * p.HW + t4));
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int item = it0 + 16 * u;
        if (item < items) {
          const int kb = item % p.kblocks, t4 = ((item / p.kblocks) * 4 + la) * 4;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int tl = t4 + k;
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = k == 0 ? v[u][j].x : (k == 1 ? v[u][j].y : (k == 2 ? v[u][j].z : v[u][j].w));
            uint32_t hw_[4], lw_[4];
            float ss = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float y0 = y[2 * q], y1 = y[2 * q + 1];
              ss = fmaf(y0, y0, ss); ss = fmaf(y1, y1, ss);
              hw_[q] = pack_f16x2(y0, y1);
              const float d0 = f16_minus_f32(hw_[q] & 0xffffu, y0), d1 = f16_minus_f32(hw_[q] >> 16, y1);
              lw_[q] = pack_f16x2(d0, d1) ^ 0x80008000u;
            }
            const uint32_t off = (uint32_t)(kb * VQ_A_KB + tl * 128 + ((((uint32_t)lb) ^ ((uint32_t)tl & 7u)) << 4));
            sts128(smem_u32(a_hi) + off, make_uint4(hw_[0], hw_[1], hw_[2], hw_[3]));
            sts128(smem_u32(a_lo) + off, make_uint4(lw_[0], lw_[1], lw_[2], lw_[3]));
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            ss += __shfl_xor_sync(0xffffffffu, ss, 4);
            if (lb == 0) z2p[kb * 128 + tl] = ss;           // one writer per (k-block, token)
          }
        }
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of the A tile -> tensor core
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                               // both CTAs: A tiles written, barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 6 + 1] = clock64();

  // ---- phase 1
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int c = 0; c < p.nchunks; ++c) {
      const int ch = (c + cl) % p.nchunks;        // clusters walk the codebook in rotated order: no L2 line is hit by all at once
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait<200>(smem_u32(empty + stage), phase ^ 1, aborted); if (aborted) goto role_done;
        if (elect_one()) {
          const uint32_t sb = smem_u32(ring + stage * VQ_STAGE);
          if (rank == 0) mbar_expect_tx(smem_u32(full + stage), (uint32_t)(2 * VQ_STAGE));
          const uint32_t fb = map_to_cta(smem_u32(full + stage), 0u);
          tma_load_3d_pair(sb, rank == 0 ? &tmB_hi : &tmB_lo, fb, kb * 64, ch * 128, 0);
          tma_load_3d_pair(sb + VQ_BX, &tmB_half, fb, kb * 64, ch * 128 + (int)rank * 64, 0);
        }
        __syncwarp();
        if (++stage == VQ_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t pdesc2 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      constexpr uint32_t pdesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0, slot = 0;
      uint32_t phase = 0, slot_phase = 0;
      for (int ch = 0; ch < p.nchunks; ++ch) {
        mbar_wait_cl(smem_u32(cempty + slot), slot_phase ^ 1, aborted); if (aborted) goto role_done;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait_cl(smem_u32(full + stage), phase, aborted); if (aborted) goto role_done;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(slot * 256);
            const uint32_t ah = desc_lo(smem_u32(a_hi) + kb * VQ_A_KB), al = desc_lo(smem_u32(a_lo) + kb * VQ_A_KB);
            const uint32_t sb = smem_u32(ring + stage * VQ_STAGE);
            const uint32_t bx = desc_lo(sb), by = desc_lo(sb + VQ_BX);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              tc_mma_f16_pair_w(d_tmem, ah + 2 * k, DESC_HI_SW128, bx + 2 * k, DESC_HI_SW128, pdesc2, (kb > 0 || k > 0) ? 1u : 0u);
              tc_mma_f16_pair_w(d_tmem + 128, al + 2 * k, DESC_HI_SW128, by + 2 * k, DESC_HI_SW128, pdesc, 1u);
            }
            tc_commit_pair(smem_u32(empty + stage), (uint16_t)3);
            if (kb == p.kblocks - 1) tc_commit_pair(smem_u32(cfull + slot), (uint16_t)3);
          }
          __syncwarp();
          if (++stage == VQ_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++slot == 2) { slot = 0; slot_phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    const int lg = warp & 3;                                  // TMEM lane quadrant of this warp
    const int tl = lg * 32 + lane;                            // token row of the tile
    float z2 = 0.f;
    for (int kb = 0; kb < p.kblocks; ++kb) z2 += z2p[kb * 128 + tl];
    const float wsi = __ldg(p.wscale_inv);
    const uint32_t cempty_leader = map_to_cta(smem_u32(cempty), 0u);
    float best = INFINITY, dsum = 0.f;
    int bi = 0x7fffffff, slot = 0;
    uint32_t slot_phase = 0;
    for (int c = 0; c < p.nchunks; ++c) {
      const int ch = (c + cl) % p.nchunks;        // same rotated order as the TMA producer
      mbar_wait<100>(smem_u32(cfull + slot), slot_phase, aborted); if (aborted) goto role_done;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(slot * 256);
      float cbest = INFINITY;
      int cbi = 0;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t r0[16], r1[16];
        tmem_ld16(taddr + c0, r0);
        tmem_ld16(taddr + 128 + c0, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int code = ch * 128 + c0 + j;
          const float dot = (__uint_as_float(r0[j]) + __uint_as_float(r1[j])) * wsi;
          const float d = (z2 + e2s[code]) - 2.f * dot;       // the reference's operation order (vqgan_arch.py:40-41)
          dsum += d;
          if (d < cbest) { cbest = d; cbi = code; }           // ascending codes inside the chunk, strict <
        }
      }
      if (cbest < best || (cbest == best && cbi < bi)) { best = cbest; bi = cbi; }   // lowest index on ties = torch.argmin, any chunk order
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(cempty_leader + (uint32_t)(slot * 8));
      if (++slot == 2) { slot = 0; slot_phase ^= 1; }
    }
    bi = min(max(bi, 0), p.K - 1);
    bidx[tl] = bi;
    const int64_t tok = (int64_t)n * p.HW + hw0 + tl;
    p.idx[tok] = (int64_t)bi;
    atomicAdd(p.hist + bi, 1u);
    double ds = (double)dsum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
    if (lane == 0) red[(warp - 2) * 2 + 1] = ds;
  }
role_done:
  if (aborted && lane == 0) s_flags[0] = 1;
  if (p.dbg && lane == 0 && warp < 6) p.dbg[blockIdx.x * 6 + 2 + (warp >= 2 ? 1 : 0)] = clock64();     // [2] producers, [3] epilogue done
  tc_fence_before();
  __syncthreads();
  // ---- phase 2: straight-through z_q (NCHW) and the squared error, all 256 threads
  if (!s_flags[0]) {
    // same item shape as phase 0: (4 consecutive tokens) x (8 channels); 16-byte loads / stores along the token axis
    const int la = lane >> 3, lb = lane & 7;
    const int items = 8 * p.kblocks;
    const float* zn = p.z + (int64_t)n * p.D * p.HW + hw0;
    float* qn = p.zq + (int64_t)n * p.D * p.HW + hw0;
    double se = 0.0;
    for (int item = warp; item < items; item += VQ_THREADS / 32) {
      const int kb = item % p.kblocks, t4 = ((item / p.kblocks) * 4 + la) * 4;
      const int c0 = kb * 64 + lb * 8;
      float4 zz[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) zz[j] = __ldg(reinterpret_cast<const float4*>(zn + (int64_t)(c0 + j) * p.HW + t4));
      float ev[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* er = p.codebook + (int64_t)bidx[t4 + k] * p.D + c0;
        const float4 e0 = __ldg(reinterpret_cast<const float4*>(er)), e1 = __ldg(reinterpret_cast<const float4*>(er + 4));
        ev[k][0] = e0.x; ev[k][1] = e0.y; ev[k][2] = e0.z; ev[k][3] = e0.w; ev[k][4] = e1.x; ev[k][5] = e1.y; ev[k][6] = e1.z; ev[k][7] = e1.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 o;
        float diff;
        diff = ev[0][j] - zz[j].x; se += (double)(diff * diff); o.x = zz[j].x + diff;      // z + (z_q - z), vqgan_arch.py:57
        diff = ev[1][j] - zz[j].y; se += (double)(diff * diff); o.y = zz[j].y + diff;
        diff = ev[2][j] - zz[j].z; se += (double)(diff * diff); o.z = zz[j].z + diff;
        diff = ev[3][j] - zz[j].w; se += (double)(diff * diff); o.w = zz[j].w + diff;
        *reinterpret_cast<float4*>(qn + (int64_t)(c0 + j) * p.HW + t4) = o;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    __shared__ double se_w[VQ_THREADS / 32];
    if (lane == 0) se_w[warp] = se;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int i = 0; i < VQ_THREADS / 32; ++i) a += se_w[i];
      for (int i = 0; i < 4; ++i) b += red[i * 2 + 1];
      p.part[(int64_t)blockIdx.x * 2] = a;
      p.part[(int64_t)blockIdx.x * 2 + 1] = b;
      __threadfence();
      s_flags[1] = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_flags[1]) {
      // last CTA of the grid: statistics (vqgan_arch.py:42,55,60-61) from the partial sums and the histogram, then clear both
      __threadfence();
      __shared__ double sred[3][VQ_THREADS];
      const int T = p.N * p.HW;
      double ent = 0.0, s_se = 0.0, s_d = 0.0;
      for (int k = threadIdx.x; k < p.K; k += VQ_THREADS) {
        const float em = (float)__ldcg(p.hist + k) / (float)T;
        ent += (double)(em * logf(em + 1e-10f));
        p.hist[k] = 0u;
      }
      for (int i = threadIdx.x; i < (int)gridDim.x; i += VQ_THREADS) { s_se += __ldcg(p.part + 2 * i); s_d += __ldcg(p.part + 2 * i + 1); }
      sred[0][threadIdx.x] = ent; sred[1][threadIdx.x] = s_se; sred[2][threadIdx.x] = s_d;
      __syncthreads();
      if (threadIdx.x == 0) {
        double e = 0.0, s2 = 0.0, d2 = 0.0;
        for (int i = 0; i < VQ_THREADS; ++i) { e += sred[0][i]; s2 += sred[1][i]; d2 += sred[2][i]; }      // fixed order
        const float mse = (float)(s2 / ((double)T * p.D));
        p.stats[0] = mse + p.beta * mse;
        p.stats[1] = expf(-(float)e);
        p.stats[2] = (float)(d2 / ((double)T * p.K));
        p.stats[3] = 0.f;
        *p.ticket = 0u;
      }
    }
  }
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 6 + 4] = clock64();
  if (aborted) {
    const long long t0 = clock64();
    while (clock64() - t0 < 400000) {}
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (p.dbg && threadIdx.x == 32) p.dbg[blockIdx.x * 6 + 5] = clock64();
}

bool vq_fused_supported(int N, int D, int HW, int K) {
  return halo_enabled() && pair_enabled() && N >= 0 && D % 64 == 0 && D >= 64 && D <= 256 && HW % 256 == 0 && K % 128 == 0 && K <= 1024;
}

int vq_fused(const float* z, const float* codebook, const void* whi, const void* wlo, const float* wscale_inv, const float* e2,
             unsigned* hist, unsigned* ticket, double* part, int N, int D, int HW, int K, float beta, float* zq, int64_t* idx,
             float* stats, cudaStream_t st, long long* dbg) {
  CFB_REQUIRE(vq_fused_supported(N, D, HW, K), "vq_fused: shape not supported");
  if (N == 0) return 0;
  TcMaps mp;
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)K, 1};
    const uint64_t str[2] = {(uint64_t)D * 2, (uint64_t)K * D * 2};
    const uint32_t box[3] = {64, 128, 1}, hbox[3] = {64, 64, 1};
    CFB_CHECK(make_map(&mp.b_hi, whi, 3, dims, str, box));
    CFB_CHECK(make_map(&mp.b_lo, wlo, 3, dims, str, box));
    CFB_CHECK(make_map(&mp.b_half, whi, 3, dims, str, hbox));
  }
  VqParams p;
  p.z = z; p.codebook = codebook; p.e2 = e2; p.wscale_inv = wscale_inv; p.zq = zq; p.idx = idx; p.stats = stats; p.part = part;
  p.hist = hist; p.ticket = ticket; p.N = N; p.D = D; p.HW = HW; p.K = K; p.kblocks = D / 64; p.nchunks = K / 128; p.beta = beta;
  p.dbg = dbg;
  constexpr int SMEM = 8 * VQ_A_KB + VQ_STAGES * VQ_STAGE + 1024 * 4 + 512 * 4 + 128 * 4 + 16 * 8 + 16 * 8 + 64 + 1024;
  static_assert(SMEM <= 232448, "shared memory budget");
  static std::atomic<uint64_t> attr_done{0};
  int dev = 0;
  CFB_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    CFB_CUDA(cudaFuncSetAttribute(vq_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * N * (HW / 256)));
  cfg.blockDim = dim3(VQ_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CFB_CUDA(cudaLaunchKernelEx(&cfg, vq_fused_kernel, mp.b_hi, mp.b_lo, mp.b_half, p));
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// Diagnostics: which shared-memory rows does a K-major SWIZZLE_128B UMMA descriptor read when its start address is
// NOT 1024-byte aligned (row-shifted views of one TMA-written tile) and how does the `base_offset` field enter?
// D = A_view * I, so D[m][n] reveals the (row, 16-byte chunk) of A that reached the tensor core.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const int* __restrict__ cfg,
                  int ncfg, int rowsA, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                        // rowsA x 128 B (<= 32 KB)
  uint8_t* sB = smem + 32768;                // 64 x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bars), 1);
    mbar_init(smem_u32(bars + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(bars), (uint32_t)(rowsA * 128 + 64 * 128));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sA)), "l"(&tmA), "r"(smem_u32(bars)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sB)), "l"(&tmB), "r"(smem_u32(bars)), "r"(0), "r"(0) : "memory");
  }
  mbar_wait(smem_u32(bars), 0);
  tc_fence_after();
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  for (int c = 0; c < ncfg; ++c) {
    const int shift = cfg[c * 3 + 0], boff = cfg[c * 3 + 1], sbo = cfg[c * 3 + 2];
    if (threadIdx.x == 0) {
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_addr = smem_u32(sA) + shift * 128 + k * 32;
        const uint64_t a_desc = (uint64_t)((a_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
                                ((uint64_t)1 << 46) | ((uint64_t)(boff & 7) << 49) | ((uint64_t)2 << 61);
        tc_mma_f16(tmem_base, a_desc, umma_desc_sw128(smem_u32(sB) + k * 32), idesc, k > 0 ? 1u : 0u);
      }
      tc_commit(smem_u32(bars + 1));
    }
    mbar_wait(smem_u32(bars + 1), (uint32_t)(c & 1));
    tc_fence_after();
    uint32_t r0[32], r1[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    tmem_ld32(taddr, r0);
    tmem_ld32(taddr + 32, r1);
    tmem_ld_wait();
    float* o = out + ((int64_t)c * 128 + warp * 32 + lane) * 64;
    for (int j = 0; j < 32; ++j) { o[j] = __uint_as_float(r0[j]); o[32 + j] = __uint_as_float(r1[j]); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
}

int umma_probe(const void* a_f16, int rowsA, const void* b_f16, const int* cfg_dev, int ncfg, float* out, cudaStream_t st) {
  CFB_REQUIRE(rowsA >= 8 && rowsA <= 256, "umma_probe: 8..256 rows");
  CUtensorMap mA, mB;
  {
    const uint64_t dims[2] = {64, (uint64_t)rowsA};
    const uint64_t str[1] = {128};
    const uint32_t box[2] = {64, (uint32_t)rowsA};
    CFB_CHECK(make_map(&mA, a_f16, 2, dims, str, box));
  }
  {
    const uint64_t dims[2] = {64, 64};
    const uint64_t str[1] = {128};
    const uint32_t box[2] = {64, 64};
    CFB_CHECK(make_map(&mB, b_f16, 2, dims, str, box));
  }
  const int smem = 32768 + 8192 + 64 + 1024;
  umma_probe_kernel<<<1, 128, smem, st>>>(mA, mB, cfg_dev, ncfg, rowsA, out);
  CFB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// Diagnostics: sustained tcgen05.mma rate (cycles per 128 x N x 16 MMA) as a function of N and of how many distinct
// TMEM accumulators consecutive MMAs rotate over (1 = every MMA depends on the previous one's accumulator).
// Operands are whatever is in shared memory (values irrelevant).  One CTA per SM.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int N, int nacc, int reps, long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * 16384 + 4 * 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = desc_lo(smem_u32(smem)), b0 = desc_lo(smem_u32(smem + 4 * 16384));
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int i = 0; i < 12; ++i) {
          const uint32_t d = tmem_base + (uint32_t)(((r * 12 + i) % nacc) * N);
          tc_mma_f16_w(d, a0 + 2 * (i & 3) + (uint32_t)((i >> 2) * 1024), DESC_HI_SW128, b0 + 2 * (i & 3) + (uint32_t)((i >> 2) * 2048),
                       DESC_HI_SW128, idesc, 1u);
        }
      }
      tc_commit(smem_u32(bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(bar), 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int umma_rate(int N, int nacc, int reps, long long* out_dev, int ctas, cudaStream_t st) {
  CFB_REQUIRE((N == 64 || N == 128 || N == 256) && nacc >= 1 && nacc * N <= 512, "umma_rate: bad configuration");
  const int smem = 4 * 16384 + 4 * 32768 + 64 + 1024;
  CFB_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_kernel<<<ctas, 128, smem, st>>>(N, nacc, reps, out_dev);
  CFB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// Diagnostics: CTA-pair MMA (tcgen05.mma.cta_group::2, M = 256 over two SMs of a TPC).  Each CTA holds its own 128 A
// rows and HALF of the B rows; the leader issues, the commit is multicast to both CTAs.  Checks which accumulator
// columns the two B halves land in and measures the sustained rate (tools/umma_pair.py).
// ------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    umma_pair_kernel(int N, int reps, float* __restrict__ vals, long long* __restrict__ info) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // A: 128 rows x 64 fp16 (16 KB); B: N/2 rows x 64 fp16 (<= 16 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  // A = rank + 1 everywhere; B half of this CTA = 1.0 (rank 0) / 2.0 (rank 1): D[row, col] = K * (rank_of_row + 1) * b(col)
  const uint32_t av = rank == 0 ? 0x3c003c00u : 0x40004000u;   // fp16 1.0 / 2.0
  const uint32_t bv = rank == 0 ? 0x3c003c00u : 0x40004000u;
  for (int i = threadIdx.x; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = av;
  for (int i = threadIdx.x; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(smem + 16384)[i] = bv;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();          // both CTAs: operands written, TMEM allocated, barriers initialised
  long long cyc = 0;
  if (warp == 0) {
    const long long t0 = clock64();
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t a0 = desc_lo(smem_u32(smem)), b0 = desc_lo(smem_u32(smem + 16384));
      if (elect_one()) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll 1
          for (int i = 0; i < 4; ++i)
            tc_mma_f16_pair_w(tmem_base, a0 + 2 * i, DESC_HI_SW128, b0 + 2 * i, DESC_HI_SW128, idesc, (r | i) ? 1u : 0u);
        }
        tc_commit_pair(smem_u32(bar), (uint16_t)3);
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(bar), 0);
    cyc = clock64() - t0;
  }
  __syncthreads();
  tc_fence_after();
  {
    // every warp reads its 32 TMEM lanes; lane 0 of warp 0 / warp 3 reports four probe columns
    uint32_t r0[32], r1[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), r0);
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(N - 32), r1);
    tmem_ld_wait();
    if ((threadIdx.x & 31) == 0 && (warp == 0 || warp == 3)) {
      float* o = vals + ((pair * 2 + rank) * 2 + (warp == 3)) * 4;
      o[0] = __uint_as_float(r0[0]); o[1] = __uint_as_float(r0[31]); o[2] = __uint_as_float(r1[0]); o[3] = __uint_as_float(r1[31]);
    }
  }
  if (threadIdx.x == 0) { info[(pair * 2 + rank) * 2] = cyc; info[(pair * 2 + rank) * 2 + 1] = (long long)tmem_base; }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

int umma_pair(int N, int reps, float* vals_dev, long long* info_dev, int ctas, cudaStream_t st) {
  CFB_REQUIRE((N == 64 || N == 128 || N == 256) && reps >= 1 && ctas >= 2 && ctas % 2 == 0, "umma_pair: bad configuration");
  const int smem = 2 * 16384 + 1024 + 64;
  CFB_CUDA(cudaFuncSetAttribute(umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_pair_kernel<<<ctas, 128, smem, st>>>(N, reps, vals_dev, info_dev);
  CFB_LAUNCH_CHECK();
  return 0;
}

}  // namespace cfb
