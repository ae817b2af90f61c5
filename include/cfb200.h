/*
 * cfb200.h -- C ABI of libcfb200.so, the B200-native (sm_100a) replacement of CodeFormer's
 * core forward pass.
 *
 * Drop-in boundary.  The reference has no FFI on this path: the arithmetic is reached
 * through PyTorch nn.Modules looked up in a registry
 *   ARCH_REGISTRY.get('CodeFormer') / .get('VQAutoEncoder')   basicsr/utils/registry.py:62-66,79
 *   CodeFormer.forward(x, w, detach_16, code_only, adain)      basicsr/archs/codeformer_arch.py:223-280
 *   VQAutoEncoder.forward(x)                                   basicsr/archs/vqgan_arch.py:385-389
 *   VectorQuantizer.forward(z) / get_codebook_feat             basicsr/archs/vqgan_arch.py:33-84
 * and the only native-extension convention the reference has is the pybind11 `m.def` modules
 * of basicsr/ops/{dcn,fused_act,upfirdn2d}/src/*.cpp built by basicsr/setup.py:118-135.
 * This header is what a maintainer binds instead (ctypes stub in INTEGRATION.md): plain C,
 * raw device/host pointers + sizes + a cudaStream_t passed as void*, int status returns
 * (0 = ok; cfb_last_error() gives the message).  No torch types cross this boundary.
 *
 * All tensors are fp32 and contiguous.  "NCHW" tensors use the reference's layout; internal
 * activations are NHWC and never leave the library.  Every call enqueues on `stream` and
 * returns without synchronising unless stated.
 */
#ifndef CFB200_H_
#define CFB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFB_VERSION 100

typedef struct cfb_net cfb_net;

/* Constructor arguments of the reference classes.
 * CodeFormer(dim_embd, n_head, n_layers, codebook_size, latent_size, connect_list)  codeformer_arch.py:162-166
 * VQAutoEncoder(img_size, nf, ch_mult, 'nearest', res_blocks, attn_resolutions, codebook_size, emb_dim, beta)
 *                                                                                    vqgan_arch.py:328-329 */
typedef struct cfb_config {
  int32_t kind;            /* 0 = VQAutoEncoder, 1 = CodeFormer */
  int32_t img_size;        /* 512 */
  int32_t nf;              /* 64 */
  int32_t n_ch_mult;       /* 6 */
  int32_t ch_mult[8];      /* 1,2,2,4,4,8 */
  int32_t res_blocks;      /* 2 */
  int32_t n_attn_res;      /* 1 */
  int32_t attn_res[4];     /* 16 */
  int32_t codebook_size;   /* 1024 */
  int32_t emb_dim;         /* 256 */
  float   beta;            /* 0.25 */
  /* CodeFormer only */
  int32_t dim_embd;        /* 512 */
  int32_t n_head;          /* 8 */
  int32_t n_layers;        /* 9 */
  int32_t latent_size;     /* 256 */
  int32_t n_connect;       /* 4 */
  int32_t connect[6];      /* 32,64,128,256 (feature sizes of connect_list) */
} cfb_config;

/* ---- library ---- */
int         cfb_version(void);
const char* cfb_last_error(void);             /* thread-local message of the last failing call */
int         cfb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- network life cycle (replaces nn.Module construction + load_state_dict) ---- */
cfb_net* cfb_net_create(const cfb_config* cfg);                 /* NULL on error */
void     cfb_net_destroy(cfb_net* net);
/* Hand one state_dict tensor (reference key name, reference layout e.g. OIHW) as a DEVICE
 * fp32 pointer; it is copied/re-laid-out at cfb_net_prepare and need not outlive it. */
int      cfb_net_set_param(cfb_net* net, const char* name, const float* dev_ptr, int64_t numel);
int      cfb_net_prepare(cfb_net* net, void* stream);          /* errors if any key is missing */
int64_t  cfb_workspace_bytes(cfb_net* net, int32_t batch);     /* scratch needed for a forward at this batch; <0 on error */
/* engine of the dense convolutions / linears: 0 = auto (tcgen05 tensor cores wherever the shape allows, default),
 * 1 = fp32 CUDA-core implicit GEMM everywhere, 2 = tcgen05 only (unsupported shapes are an error) */
int      cfb_net_set_engine(cfb_net* net, int32_t engine);
/* parity hook: after stage `name` ("enc.<i>", "gen.<i>", "fuse.<size>", "ft.<l>", "quant") of the next forwards, copy
 * that NHWC fp32 activation to dst (device, `capacity` floats).  dst = NULL removes the hook. */
int      cfb_net_capture(cfb_net* net, const char* stage, float* dst, int64_t capacity);
/* number of kernel launches the last forward on this net enqueued (bench.py "gpu_launches") */
int64_t  cfb_last_launch_count(cfb_net* net);

/* ---- CodeFormer.forward (codeformer_arch.py:223-280) ----
 * x [B,3,512,512] NCHW in [-1,1]; out [B,3,512,512] NCHW (unclamped; may be NULL when code_only);
 * logits [B,256,K]; lq_feat [B,256,16,16] NCHW; top_idx [B,256] int64 (may be NULL).
 * All DEVICE pointers.  `w` is the fidelity weight (branch w>0, :276). */
int cfb_codeformer_forward(cfb_net* net, const float* x, float* out, float* logits, float* lq_feat,
                           int64_t* top_idx, int32_t batch, float w, int32_t adain, int32_t code_only,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* Same call with HOST buffers (pinned recommended): H2D of x, forward, D2H of out/logits/lq_feat,
 * all on `stream`, then one stream synchronise.  dev_scratch must hold the device copies:
 * cfb_host_io_bytes(net,batch) bytes, in addition to the workspace. */
int64_t cfb_host_io_bytes(cfb_net* net, int32_t batch);
int cfb_codeformer_forward_host(cfb_net* net, const float* x_host, float* out_host, float* logits_host,
                                float* lq_feat_host, int32_t batch, float w, int32_t adain,
                                void* dev_scratch, int64_t dev_scratch_bytes,
                                void* workspace, int64_t workspace_bytes, void* stream);

/* f1: the forward with the caller's image plumbing fused into the first and last conv.
 * faces_bgr / restored_bgr: DEVICE uint8 [batch,512,512,3] HWC BGR, i.e. face_helper.cropped_faces as they are
 * (inference_codeformer.py:197).  Replaces img2tensor(face/255.) + normalize(0.5,0.5) -> net(x,w,adain)[0] ->
 * tensor2img(rgb2bgr, min_max=(-1,1)).astype(uint8)   (inference_codeformer.py:199-213, basicsr/utils/img_util.py:9-35,38-94)
 * with the same fp32 arithmetic and rounding (round-half-even).  logits / lq_feat / top_idx are optional. */
int cfb_codeformer_forward_u8(cfb_net* net, const uint8_t* faces_bgr, uint8_t* restored_bgr, float* logits, float* lq_feat,
                              int64_t* top_idx, int32_t batch, float w, int32_t adain,
                              void* workspace, int64_t workspace_bytes, void* stream);
/* the same with HOST uint8 buffers (0.79 MB per face each way instead of 3.1 MB): H2D, forward, D2H, stream sync.
 * dev_scratch >= cfb_host_io_bytes(net, batch). */
int cfb_codeformer_restore_host(cfb_net* net, const uint8_t* faces_host, uint8_t* restored_host, int32_t batch, float w,
                                int32_t adain, void* dev_scratch, int64_t dev_scratch_bytes,
                                void* workspace, int64_t workspace_bytes, void* stream);
/* the plumbing alone (unit parity): uint8 HWC BGR [n,hw,3] <-> fp32 NCHW RGB [n,3,hw] in [-1,1] */
int cfb_u8_to_input(const uint8_t* img_bgr_hwc, float* x_nchw, int32_t n, int32_t hw, void* stream);
int cfb_output_to_u8(const float* x_nchw, uint8_t* img_bgr_hwc, int32_t n, int32_t hw, void* stream);
/* ---- VQAutoEncoder.forward (vqgan_arch.py:385-389) ----
 * out [B,3,512,512]; idx [B*256] int64; stats[4] = {codebook_loss, perplexity, mean_distance, 0};
 * min_encodings [B*256,K] one-hot fp32 or NULL (materialised only when asked). */
int cfb_vqae_forward(cfb_net* net, const float* x, float* out, int64_t* idx, float* stats,
                     float* min_encodings, int32_t batch,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* ---- VectorQuantizer.forward (vqgan_arch.py:33-70), standalone (config 3 microbench) ----
 * z [B,D,H,W] NCHW, codebook [K,D]; z_q [B,D,H,W] NCHW (= z + (E[idx]-z), :57); idx [B*H*W] int64;
 * stats[4] as above; min_encodings optional.  workspace >= cfb_vq_workspace_bytes. */
int64_t cfb_vq_workspace_bytes(int32_t batch, int32_t hw, int32_t dim, int32_t codes);
int cfb_vq_nearest(const float* z, const float* codebook, int32_t batch, int32_t h, int32_t w,
                   int32_t dim, int32_t codes, float beta, float* z_q, int64_t* idx, float* stats,
                   float* min_encodings, void* workspace, int64_t workspace_bytes, void* stream);

/* VectorQuantizer.forward, fused path (vqgan_arch.py:33-70; BASELINE configs[2]): the codebook is split for the tensor cores and
 * its |e|^2 computed ONCE (cfb_vq_prepare, redo when the embedding changes).  A call is then ONE kernel on the caller's NCHW
 * tensors when h*w % 256 == 0, dim <= 256, codes <= 1024 (vq_fused_kernel: z tile -> shared-memory operand planes, codebook
 * through TMA, distances on tcgen05 with the argmin read out of TMEM, gather + straight-through z_q + loss, statistics by the
 * last CTA; the prepared buffer also holds the kernel's self-cleaning histogram, so calls sharing one prepared buffer must
 * not overlap), otherwise 4 launches (operand planes + |z|^2, distance GEMM with the argmin in its epilogue, candidate
 * reduction + gather, statistics).  The [tokens, codes] matrix is never stored.  Same outputs as cfb_vq_nearest.
 * cfb_vq_fast_supported: 16x16-style latents (h*w % 128 == 0), dim % 64 == 0, codes % 128 == 0, sm_100. */
int32_t cfb_vq_fast_supported(int32_t batch, int32_t h, int32_t w, int32_t dim, int32_t codes);
int64_t cfb_vq_prepared_bytes(int32_t codes, int32_t dim);
int     cfb_vq_prepare(const float* codebook, int32_t codes, int32_t dim, void* prepared, int64_t prepared_bytes, void* stream);
int64_t cfb_vq_fast_workspace_bytes(int32_t batch, int32_t hw, int32_t dim, int32_t codes);
int     cfb_vq_nearest_fast(const float* z, const float* codebook, const void* prepared, int32_t batch, int32_t h, int32_t w,
                            int32_t dim, int32_t codes, float beta, float* z_q, int64_t* idx, float* stats, float* min_encodings,
                            void* workspace, int64_t workspace_bytes, void* stream);
/* ---- VectorQuantizer.get_codebook_feat (vqgan_arch.py:72-84) ---- idx [n] int64 -> z_q [B,D,H,W] NCHW */
int cfb_codebook_lookup(const int64_t* idx, const float* codebook, int32_t batch, int32_t h, int32_t w,
                        int32_t dim, int32_t codes, float* z_q, void* stream);

/* ---- per-kernel entry points (unit-parity tests; NHWC fp32 device tensors) ---- */
/* conv2d: in [N,H,W,Cin] NHWC, weight OIHW [Cout,Cin,k,k] (reference layout), bias [Cout] or NULL.
 * mode: 0 = 'same' k=1|3 stride 1 (nn.Conv2d padding=k/2); 1 = Downsample (pad right/bottom 1, 3x3 s2 p0,
 * vqgan_arch.py:122-126); 2 = Upsample (nearest x2 then 3x3 p1, vqgan_arch.py:134-138).
 * in_scale/in_shift [N,Cin] optional fused per-sample affine (GroupNorm), in_act 0|1(SiLU);
 * residual [N,Ho,Wo,Cout] optional; out_act 0 | 1 LeakyReLU(0.2) | 2 GELU(erf).
 * engine: 0 = auto, 1 = fp32 CUDA-core implicit GEMM, 2 = tcgen05 split-fp16 tensor-core implicit GEMM. */
int cfb_conv2d_nhwc(const float* in, const float* weight_oihw, const float* bias, float* out,
                    int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t ksize, int32_t mode,
                    const float* in_scale, const float* in_shift, int32_t in_act,
                    const float* residual, int32_t out_act, int32_t engine,
                    void* workspace, int64_t workspace_bytes, void* stream);
int64_t cfb_conv2d_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t ksize, int32_t mode);
/* GroupNorm(32,C,eps) statistics folded with the affine: scale[n,c] = rstd*gamma, shift[n,c] = beta - mean*rstd*gamma
 * (vqgan_arch.py:14-15).  x [N,HW,C] NHWC.  workspace >= cfb_gn_workspace_bytes. */
int64_t cfb_gn_workspace_bytes(int32_t n, int32_t hw, int32_t c);
int cfb_group_norm_coef(const float* x, const float* gamma, const float* beta, float* scale, float* shift,
                        int32_t n, int32_t hw, int32_t c, int32_t groups, float eps,
                        void* workspace, int64_t workspace_bytes, void* stream);
/* y = act(x*scale[n,c]+shift[n,c]) materialised (NHWC) */
int cfb_affine_act(const float* x, const float* scale, const float* shift, float* y,
                   int32_t n, int32_t hw, int32_t c, int32_t act, void* stream);
/* softmax(q k^T * scale) v for `heads` heads of width d packed in rows of pitch (in floats); S tokens per batch */
int cfb_attention(const float* q, const float* k, const float* v, float* out,
                  int32_t batch, int32_t tokens, int32_t heads, int32_t d,
                  int32_t q_pitch, int32_t k_pitch, int32_t v_pitch, int32_t o_pitch, float scale, void* stream);
/* LayerNorm over the last dim (eps 1e-5); y2 (optional) = y + pos[t % pos_rows] */
int cfb_layer_norm(const float* x, const float* gamma, const float* beta, float* y, float* y2,
                   const float* pos, int32_t pos_rows, int32_t rows, int32_t c, void* stream);
/* AdaIN of codeformer_arch.py:29-43 on NHWC [B,HW,C] */
int cfb_adain_nhwc(const float* content, const float* style, float* out, int32_t batch, int32_t hw, int32_t c, void* stream);
/* diagnostics: D = A_view * I for row-shifted 128B-swizzled UMMA descriptor views (tools/umma_probe.py).
 * a_f16 [rows_a,64] fp16, b_f16 [64,64] fp16, cfg_dev [ncfg][3] = {shift_rows, base_offset, sbo_bytes}, out [ncfg,128,64] */
int cfb_debug_umma_probe(const void* a_f16, int32_t rows_a, const void* b_f16, const int32_t* cfg_dev, int32_t ncfg,
                         float* out, void* stream);
/* diagnostics: sustained tcgen05.mma issue rate; out_dev[ctas] receives the cycles for reps*12 MMAs of 128 x n x 16
 * rotating over `nacc` TMEM accumulators (tools/umma_rate.py) */
/* CTA-pair probe (tcgen05.mma.cta_group::2): vals [ctas][2][4] accumulator samples, info [ctas][2] = cycles, tmem base */
int cfb_debug_umma_pair(int32_t n, int32_t reps, float* vals_dev, int64_t* info_dev, int32_t ctas, void* stream);
int cfb_debug_umma_rate(int32_t n, int32_t nacc, int32_t reps, int64_t* out_dev, int32_t ctas, void* stream);
/* diagnostics / bench: average device time (CUDA events on `stream`) of the tcgen05 conv KERNEL alone over `reps` launches
 * (bench.py roofline).  With in_scale/in_shift (per-(n,cin) GroupNorm affine) and in_act the kernel is the one the forward
 * launches for a GroupNorm+SiLU consumer: the fused-operand-transform variant reading the fp32 activation (all-in, no prep
 * pass exists); without them the weights are split and the raw operand planes prepared once outside the timed region. */
int cfb_debug_time_conv(const float* in, const float* weight_oihw, float* out, int32_t n, int32_t h, int32_t w, int32_t cin,
                        int32_t cout, int32_t ksize, int32_t mode, int32_t reps, void* workspace, int64_t workspace_bytes,
                        void* stream, const float* in_scale, const float* in_shift, int32_t in_act, float* ms_per_launch);
/* ---- RRDBNet: the upsampler behind RealESRGANer.enhance ----
 * /root/reference/basicsr/archs/rrdbnet_arch.py:67-120 (constructor :86, forward :103-119); the caller's tiling loop
 * (basicsr/utils/realesrgan_utils.py:100-175) stays in Python (codeformer_b200/upsampler.py).
 * Parameters are the reference's state-dict names (conv_first, body.{i}.rdb{1,2,3}.conv{1..5}, conv_body, conv_up1, conv_up2,
 * conv_hr, conv_last; .weight OIHW / .bias).  x: [batch, num_in_ch, h, w] fp32 NCHW, any h, w (multiples of 2 for scale 2, of 4
 * for scale 1: pixel_unshuffle); out: [batch, num_out_ch, h*scale, w*scale].  Built for num_feat = 64, num_grow_ch = 32. */
typedef struct cfb_rrdb cfb_rrdb;
cfb_rrdb* cfb_rrdb_create(int32_t num_in_ch, int32_t num_out_ch, int32_t scale, int32_t num_feat, int32_t num_block, int32_t num_grow_ch);
void      cfb_rrdb_destroy(cfb_rrdb* net);
int       cfb_rrdb_set_param(cfb_rrdb* net, const char* name, const float* dev_ptr, int64_t numel);
int       cfb_rrdb_prepare(cfb_rrdb* net, void* stream);
int64_t   cfb_rrdb_workspace_bytes(cfb_rrdb* net, int32_t batch, int32_t h, int32_t w);
int       cfb_rrdb_forward(cfb_rrdb* net, const float* x, float* out, int32_t batch, int32_t h, int32_t w,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* ---- ParseNet: face parsing of the restored face for the paste-back blend ----
 * /root/reference/facelib/parsing/parsenet.py:140-194; built by init_parsing_model('parsenet') as ParseNet(in_size=512,
 * out_size=512, parsing_ch=19) (facelib/parsing/__init__.py:13) and called at facelib/utils/face_restoration_helper.py:457-462.
 * Parameters are the reference's state-dict names (encoder.0.conv2d, {encoder,body,decoder}.{i}.{shortcut_func,conv1,conv2}.
 * conv2d.{weight,bias}, ...norm.norm.{weight,bias,running_mean,running_var}); eval-mode BatchNorm is folded at prepare.
 * x: [batch,3,h,w] fp32 NCHW in [-1,1]; out_mask: [batch, parsing_ch, h, w] logits; out_img (optional): [batch,3,h,w].
 * cfb_parse_argmax: classes = out_mask.argmax(1) (first maximum) and/or the caller's 0/255 face mask of
 * face_restoration_helper.py:463-468 (MASK_COLORMAP), both uint8 [batch, h*w]. */
typedef struct cfb_parsenet cfb_parsenet;
cfb_parsenet* cfb_parsenet_create(int32_t in_size, int32_t out_size, int32_t min_feat_size, int32_t base_ch, int32_t parsing_ch,
                                  int32_t res_depth, int32_t ch_min, int32_t ch_max);
void      cfb_parsenet_destroy(cfb_parsenet* net);
int       cfb_parsenet_set_param(cfb_parsenet* net, const char* name, const float* dev_ptr, int64_t numel);
int       cfb_parsenet_prepare(cfb_parsenet* net, void* stream);
int64_t   cfb_parsenet_workspace_bytes(cfb_parsenet* net, int32_t batch, int32_t h, int32_t w);
int       cfb_parsenet_forward(cfb_parsenet* net, const float* x, float* out_mask, float* out_img, int32_t batch, int32_t h, int32_t w,
                               void* workspace, int64_t workspace_bytes, void* stream);
int       cfb_parse_argmax(const float* logits_nchw, uint8_t* classes, uint8_t* mask, int32_t batch, int32_t channels, int64_t hw,
                           void* stream);

/* One 3x3 conv of the generalised fused-transform engine (the building block of RRDBNet / ParseNet; test entry point).
 * in: NHWC buffer of in_pitch channels per pixel, channels [0, cin) are read; any h x w.  upsample != 0: nearest x2 first.
 * pad_mode 0 zero / 1 reflect / 2 replicate (of the low-resolution tensor when upsampling).  subsample != 0: stride 2 (the
 * even output positions of the stride-1 result; h, w even).  out: NHWC buffer of out_pitch channels, the cout real channels
 * go to [out_c0, out_c0 + cout).  out = lrelu?(conv + bias + residual) * post_scale + residual2 (post only with residual2). */
int64_t cfb_conv2d_gen_workspace_bytes(int32_t cin, int32_t cout);
int cfb_conv2d_gen_nhwc(const float* in, int32_t in_pitch, const float* weight_oihw, const float* bias, float* out,
                        int32_t out_pitch, int32_t out_c0, int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout,
                        int32_t upsample, int32_t pad_mode, int32_t subsample, int32_t out_act, const float* residual,
                        int32_t res_pitch, const float* residual2, int32_t res2_pitch, float post_scale, void* workspace,
                        int64_t workspace_bytes, void* stream);

/* Asynchronous failures.  Kernels never trap and never leave a sticky CUDA error behind (the reference's callers catch
 * RuntimeError and fall back to the input face, inference_codeformer.py:209-211; web-demos/hugging_face/app.py:176): a
 * barrier time-out of the tensor-core pipeline or an activation outside the fp16 operand range (|x| > 65504) sets a bit
 * in a host-mapped status word.  It is reported (status 1 + cfb_last_error) by the NEXT forward on that device, by the
 * *_host entry points right after their stream synchronisation, and by this call -- use it after synchronising the stream
 * of an asynchronous forward.  The context stays usable; the reporting call clears the condition. */
int cfb_check_async_status(void);
/* test hooks: barrier time-out in SM cycles (default 4e9, about 2 s); kind != 0 makes the next tcgen05 conv launch drop one
 * TMA load so that its pipeline times out (tests/test_gpu_faults.py) */
int cfb_debug_set_wait_limit(int64_t cycles);
int cfb_debug_inject_fault(int32_t kind);
/* diagnostics (tools/tc_stamps.py): while `stamps` (device memory, >= 32 int64) is non-NULL, CTA 0 of every tcgen05 conv launch
 * writes the SM cycle counter at its role hand-offs ([0] entry, [1] set-up done, [2] first TMA, [3] first MMA, [4] last MMA
 * issued, [5]/[6] first/last accumulator seen by the epilogue, [7] epilogue K loop done, [13] epilogue done, [8]/[9] tear-down,
 * [10..12] transform roles, [14]/[15] %globaltimer at entry / exit).  NULL switches it off.  The stamps exist only in builds of the
 * library with -DCFB_TC_STAMPS=1 (they cost 6-13 % of every conv kernel); the production build accepts the call and ignores it. */
int cfb_debug_set_stamps(int64_t* stamps);
/* layout plumbing */
int cfb_nchw_to_nhwc(const float* in, float* out, int32_t n, int32_t c, int32_t hw, void* stream);
int cfb_nhwc_to_nchw(const float* in, float* out, int32_t n, int32_t c, int32_t hw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFB200_H_ */
