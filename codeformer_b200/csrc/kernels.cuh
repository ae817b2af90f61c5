// Internal launcher interface between the host runtime (runtime.cu / capi.cu) and the kernels.
// Every launcher enqueues on `st` and returns the launch status; none synchronises.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace cfb {

// ---- error plumbing --------------------------------------------------------------------------
void set_error(const std::string& msg);
#define CFB_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::cfb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                       std::to_string(__LINE__));                                                 \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)
#define CFB_LAUNCH_CHECK()                                                                        \
  do {                                                                                            \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess) {                                                                      \
      ::cfb::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) + " @" +    \
                       __FILE__ + ":" + std::to_string(__LINE__));                                \
      return 1;                                                                                   \
    }                                                                                             \
    ::cfb::count_launch();                                                                        \
  } while (0)
#define CFB_CHECK(x)                                                                              \
  do {                                                                                            \
    if ((x) != 0) return 1;                                                                       \
  } while (0)
#define CFB_REQUIRE(cond, msg)                                                                    \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      ::cfb::set_error(std::string(msg) + " [" #cond "] @" + __FILE__ + ":" + std::to_string(__LINE__)); \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

// asynchronous device status word (host-mapped): kernels report a barrier time-out or an fp16 operand overflow here instead
// of trapping; the host turns it into an error without poisoning the context (runtime.cu)
#define CFB_STATUS_TIMEOUT 1u
#define CFB_STATUS_OVERFLOW 2u
int async_status_init(cudaStream_t st);      // per-device one-time setup (idempotent)
int async_status_check(const char* where);   // non-zero + set_error() when a kernel reported a failure since the last check
int tc_bind_status_word(unsigned* host_mapped_dev_ptr, long long wait_limit_cycles);   // conv_tc.cu: device symbols of this device
int tc_clear_abort();
int tc_set_stamps(long long* dev_ptr);     // diagnostics: phase stamps of CTA 0 (cfb_debug_set_stamps)
int tc_inject_fault(int kind);               // test hook: the next tcgen05 conv launch drops one TMA load (-> barrier time-out)

void count_launch();
int64_t launch_count();
void reset_launch_count();

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// The forward is a chain of ~340 dependent launches; with plain stream order each one pays the launch latency and its own
// set-up (barrier init, TMEM allocation, tensor-map prefetch: 3-5 us measured, tools/tc_stamps.py) after the previous grid has
// drained.  Kernels launched through CFB_LAUNCH_PDL carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may
// start as soon as every CTA of the previous grid has executed pdl_launch_dependents() (SM resources permitting), do their
// set-up, and block in pdl_wait() until the previous grid has COMPLETED and its memory is visible.  Rules kept everywhere:
//   * every thread calls pdl_wait() before its first access to global memory another kernel may have written (and before its
//     own first global write), and before any early return;
//   * a kernel without the attribute behaves as before (the instructions are no-ops there), so the two kinds mix freely.
// CFB_PDL=0 switches the attribute off (A/B timing).
bool pdl_enabled();
#ifdef __CUDACC__
#ifndef CFB_PDL_DEVICE
#define CFB_PDL_DEVICE 1
#endif
#if CFB_PDL_DEVICE
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else      // A/B builds without the device side (the host side then never sets the attribute: pdl_enabled() is false)
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_launch_dependents() {}
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// launch + error plumbing + launch counter (kernel must call pdl_wait())
#define CFB_LAUNCH_PDL(kernel, grid, block, smem, st, ...)                                          \
  do {                                                                                              \
    cudaError_t _e = ::cfb::launch_pdl(kernel, grid, block, smem, st, __VA_ARGS__);                 \
    if (_e != cudaSuccess) {                                                                        \
      ::cfb::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) + " @" +      \
                       __FILE__ + ":" + std::to_string(__LINE__));                                  \
      return 1;                                                                                     \
    }                                                                                               \
    ::cfb::count_launch();                                                                          \
  } while (0)
#endif

enum InAct { IN_NONE = 0, IN_SILU = 1 };
enum OutAct { OUT_NONE = 0, OUT_LRELU = 1, OUT_GELU = 2 };
enum ConvMode { CONV_SAME = 0, CONV_DOWN = 1, CONV_UP = 2 };

// Convolution / linear layer as an implicit GEMM over NHWC fp32 activations.
//   M = N*Ho*Wo output pixels (tokens), N = Cout, K = taps*Cin.
struct ConvArgs {
  const float* in = nullptr;        // [N,H,W,Cin]  (H,W are the stored dims; CONV_UP reads (y>>1,x>>1))
  int N = 0, H = 0, W = 0, Cin = 0;
  int Ho = 0, Wo = 0, Cout = 0;
  int ksize = 3;                    // 1 | 3
  int mode = CONV_SAME;
  const float* wgt_f32 = nullptr;   // [taps][Cin][Cout] fp32 (CUDA-core engine)
  const void* wgt_hi = nullptr;     // [taps][Cout][Cin] fp16 hi   (tensor-core engine)
  const void* wgt_lo = nullptr;     // [taps][Cout][Cin] fp16 lo
  const float* wscale_inv = nullptr; // device scalar 2^-k undoing the power-of-two scaling of the fp16 weight split
  const float* bias = nullptr;      // [Cout] | null
  const float* in_scale = nullptr;  // [N,Cin] fused per-sample affine (GroupNorm folded) | null
  const float* in_shift = nullptr;
  int in_act = IN_NONE;
  const float* residual = nullptr;  // [N,Ho,Wo,Cout] | null
  int out_act = OUT_NONE;
  // SFT epilogue (codeformer_arch.py:155-156): out = dec + w*(dec*scale + conv)
  const float* sft_dec = nullptr;
  const float* sft_scale = nullptr;
  float sft_w = 0.f;
  float* out = nullptr;             // [N,Ho,Wo,Cout]
  // tensor-core engine only: GroupNorm(32) partial sums of `out`, [N*tiles_per_image*4][32 groups][2] floats
  float* gn_part = nullptr;
  // tensor-core engine only: also emit `out` as fp16 hi/lo operand planes for a following conv that consumes it raw
  void* out_planes = nullptr;       // [hi plane | lo plane], each align1024(N*Ho*Wo*Cout*2) bytes
  bool skip_prep = false;           // operand planes in `scratch` are already valid (kernel-only timing)
  bool xform = false;      // tensor engine: read the fp32 activation `in` directly and apply in_scale/in_shift/in_act + the fp16 hi/lo
                           // split inside the conv kernel (tc_can_xform); in_scale == null: plain split of the raw values
  // VectorQuantizer distance GEMM with the argmin in the epilogue (vqgan_arch.py:40-46): instead of storing z.E^T, every
  // epilogue thread forms d = (|z|^2 + |e|^2) - 2 z.e for its token row and column slice and writes its first minimum
  const float* vq_e2 = nullptr;     // [Cout] |e_j|^2
  const float* vq_z2 = nullptr;     // [tokens] |z_t|^2
  float2* vq_cand = nullptr;        // [tokens][2 * Cout / BN] (distance, index bits)
  double* vq_dpart = nullptr;       // [m_tiles * n_tiles * 8] partial sums of all distances (mean_distance)
  bool halo1x1 = false;    // 1x1 conv with a GroupNorm-affine input (AttnBlock q,k,v): run it on the halo engine (patch = tile, one
                           // tap) so that the fused operand transform applies -- no separate operand-preparation pass
  const float* in2 = nullptr;   // xform only: channels [Cin1, Cin) come from this second NHWC tensor (torch.cat of Fuse_sft_block)
  int Cin1 = 0;
  // ---- generalised addressing (xform only; ParseNet / RRDBNet rows f3 / f4): any H x W (ragged tiles), padding mode of the
  // 3x3 window, source / destination inside wider NHWC buffers (dense blocks), stride 2 by subsampling, second residual
  bool gen = false;
  int in_pitch = 0;             // channels per pixel of the buffer `in` points into (0: Cin); Cin is the 64-aligned window read
  int pad_mode = 0;             // 0 zero, 1 reflect, 2 replicate
  bool subsample = false;       // keep the even output positions only: out is [N, Ho/2, Wo/2, ...] (Ho, Wo even)
  int out_pitch = 0, out_c0 = 0;   // destination channels per pixel (0: Cout) and channel offset
  int cout_valid = 0;           // real output channels (0: Cout); Cout itself is the 64-aligned padded count of the weights
  int res_pitch = 0;            // channels per pixel of `residual` (0: out_pitch)
  const float* residual2 = nullptr;   // out = act(conv + bias + residual) * post_scale + residual2
  int res2_pitch = 0;
  float post_scale = 1.f;
};

int conv_f32(const ConvArgs& a, cudaStream_t st);                       // CUDA-core fp32 implicit GEMM
// first conv: x NCHW [N,3,H,W] -> NHWC [N,H,W,Cout], 3x3 p1; weight [27][Cout] (tap-major, then cin)
// gn_part (optional): GroupNorm(32) partial sums of `out`, [N * H*W/32 slots][32 groups][sum, sum of squares] (the layout
// gn_coef_from_partials reads; slots per image = H*W/32)
int conv_first(const float* x_nchw, const float* wgt, const float* bias, float* out, int N, int H, int W, int Cout,
               cudaStream_t st, float* gn_part = nullptr);
// last conv: NHWC [N,H,W,Cin] (+ fused affine) -> NCHW [N,3,H,W], 3x3 p1; weight [9][Cin][3]
int conv_last(const float* in, const float* in_scale, const float* in_shift, const float* wgt, const float* bias,
              float* out_nchw, int N, int H, int W, int Cin, cudaStream_t st);
// the same two kernels with the caller's image plumbing fused in (inference_codeformer.py:199-206,
// basicsr/utils/img_util.py:9-35,38-94): uint8 HWC BGR face in, uint8 HWC BGR restored face out
int conv_first_u8(const unsigned char* x_bgr_hwc, const float* wgt, const float* bias, float* out, int N, int H, int W,
                  int Cout, cudaStream_t st, float* gn_part = nullptr);
int conv_last_u8(const float* in, const float* in_scale, const float* in_shift, const float* wgt, const float* bias,
                 unsigned char* out_bgr_hwc, int N, int H, int W, int Cin, cudaStream_t st);
int u8_to_input(const unsigned char* img_bgr_hwc, float* x_nchw, int N, int64_t HW, cudaStream_t st);
int output_to_u8(const float* x_nchw, unsigned char* img_bgr_hwc, int N, int64_t HW, cudaStream_t st);

// thin convolutions of the caller-side networks (rows f3 / f4): image channels -> 64 features and 64 features -> image channels
int conv_thin_in(const float* x_nchw, const float* wgt_tck, const float* bias, float* out, int N, int H, int W, int Cimg, int us,
                 int pad_mode, int out_pitch, int out_c0, cudaStream_t st);
int conv_thin_out(const float* in_nhwc64, const float* wgt_tcp, const float* bias, float* out_nchw, int N, int H, int W, int Cout,
                  int pad_mode, cudaStream_t st);
int relayout_thin_out(const float* oihw, float* out, int Cout, cudaStream_t st);
int fold_bn(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* wout,
            float* bout, int Cout, int per_out, cudaStream_t st);
int parse_argmax(const float* logits_nchw, unsigned char* cls, unsigned char* mask, int N, int C, int64_t HW, cudaStream_t st);
int scale_scalar(float* p, float f, cudaStream_t st);
int scale_vec(float* p, int n, float f, cudaStream_t st);

// weight re-layout: OIHW -> [taps][Cin][Cout]
int relayout_oihw_to_tck(const float* oihw, float* out, int Cout, int Cin, int k, cudaStream_t st);

// GroupNorm statistics -> per-(n,c) scale/shift.  partials: workspace of gn_partial_count() doubles
size_t gn_workspace_bytes(int N, int HW, int C);
int gn_coef(const float* x, const float* gamma, const float* beta, float* scale, float* shift, int N, int HW, int C,
            int groups, float eps, void* ws, cudaStream_t st);
// finalize from the tensor-core epilogue's partial sums (slots per image = tiles_per_image*4)
// scratch: gn_final_scratch_bytes(N, slots) bytes; counters: N zero-initialised unsigned (left zero again by the kernel)
size_t gn_final_scratch_bytes(int N, int slots);
int gn_coef_from_partials(const float* part, int slots, const float* gamma, const float* beta, float* scale, float* shift,
                          int N, int HW, int C, int groups, float eps, void* scratch, unsigned* counters, cudaStream_t st);
// partial sums of cat([a,b]) (2C channels, 32 groups) from the partial sums of a and b (C channels each)
int gn_cat_partials(const float* a_part, const float* b_part, float* out_part, int64_t total_slots, cudaStream_t st);
int affine_act(const float* x, const float* scale, const float* shift, float* y, int N, int HW, int C, int act,
               cudaStream_t st);

// out (fp32) and/or out_planes (fp16 hi | lo operand planes of the same [B*S, o_pitch] matrix) receive the result
int attention(const float* q, const float* k, const float* v, float* out, int B, int S, int heads, int d, int q_pitch,
              int k_pitch, int v_pitch, int o_pitch, float scale, cudaStream_t st, void* out_planes = nullptr);
int layer_norm(const float* x, const float* gamma, const float* beta, float* y, float* y2, const float* pos,
               int pos_rows, int rows, int C, cudaStream_t st);
int layer_norm_planes(const float* x, const float* gamma, const float* beta, void* y_planes, void* y2_planes, const float* pos,
                      int pos_rows, int rows, int C, cudaStream_t st);
int add_pos(const float* x, const float* pos, float* y, int rows, int pos_rows, int C, cudaStream_t st);
// logits [T,K] -> idx [T] (first max), quant [T,D] = E[idx]
int argmax_gather(const float* logits, const float* codebook, int64_t* idx, float* quant, int T, int K, int D,
                  cudaStream_t st);
int gather_rows(const int64_t* idx, const float* codebook, float* out, int T, int K, int D, cudaStream_t st);
int adain_nhwc(const float* content, const float* style, float* out, int B, int HW, int C, cudaStream_t st);
int nchw_to_nhwc(const float* in, float* out, int N, int C, int HW, cudaStream_t st);
int nhwc_to_nchw(const float* in, float* out, int N, int C, int HW, cudaStream_t st);
int concat_channels(const float* a, const float* b, float* out, int64_t pixels, int Ca, int Cb, cudaStream_t st);

// VectorQuantizer.forward core on token-major z [T,D] (NHWC); writes idx, zq_st = z + (E[idx]-z) [T,D],
// stats = {loss, perplexity, mean_distance, 0}; onehot optional [T,K]
size_t vq_workspace_bytes(int T, int D, int K);
int vq_nearest(const float* z, const float* codebook, int T, int D, int K, float beta, int64_t* idx, float* zq,
               float* stats, float* onehot, void* ws, cudaStream_t st);

// fused path (config 3): z NCHW -> operand planes + |z|^2 (vq_prep_nchw), distance GEMM with argmin epilogue (conv_tc),
// candidate reduction + gather + straight-through z_q in NCHW + loss partials (vq_select_cand), statistics (vq_final2)
int vq_prep_nchw(const float* z_nchw, void* planes, float* z2, unsigned* hist, int N, int D, int HW, int K, cudaStream_t st);
int vq_select_cand(const float* z_nchw, const float* codebook, const float2* cand, int ncand, int N, int D, int HW, int K,
                   int64_t* idx, float* zq_nchw, double* se_part, unsigned* hist, cudaStream_t st);
int vq_final2(const double* se_part, int n_se, const double* d_part, int n_d, const unsigned* hist, int T, int D, int K, float beta,
              float* stats, cudaStream_t st);
int vq_e2(const float* codebook, float* e2, int K, int D, cudaStream_t st);
int onehot_from_idx(const int64_t* idx, float* onehot, int T, int K, cudaStream_t st);

// tensor-core variant: dots [T,K] = z . E^T already computed (tcgen05 1x1 conv); selects the nearest code per token
size_t vq_select_workspace_bytes(int T, int K);
int vq_select_from_dots(const float* z, const float* codebook, const float* dots, int T, int D, int K, float beta, int64_t* idx,
                        float* zq, float* stats, float* onehot, void* ws, cudaStream_t st);
}  // namespace cfb
