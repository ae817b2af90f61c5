// Host runtime of libcfb200: the network plan (block lists of the reference constructors), weight
// preparation, the stream-ordered workspace arena, and the forward passes that enqueue the kernels.
//
// Mirrors, block for block:
//   Encoder / Generator block lists      /root/reference/basicsr/archs/vqgan_arch.py:229-323
//   VQAutoEncoder.forward                vqgan_arch.py:385-389
//   CodeFormer.forward                   /root/reference/basicsr/archs/codeformer_arch.py:223-280
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <array>
#include <atomic>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cfb200.h"
#include "kernels.cuh"
#include "conv_tc.cuh"

namespace cfb {

// ---- error / counters --------------------------------------------------------------------------
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const std::string& last_error() { return g_err; }
static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
#if defined(CFB_PDL_DEVICE) && !CFB_PDL_DEVICE
  return false;
#endif
  static const bool v = [] { const char* e = getenv("CFB_PDL"); return !(e && atoi(e) == 0); }();
  return v;
}
int64_t launch_count() { return g_launches.load(); }
void reset_launch_count() { g_launches.store(0); }

// ---- asynchronous device status (barrier time-out / fp16 operand overflow) ---------------------------------
// One host-mapped word per device.  Kernels never trap: they set bits here (conv_tc.cu) and the host reports them as an
// ordinary error at the next check point -- the CUDA context stays usable, the caller's per-face fallback keeps working
// (SURVEY.md section 8(b) "Errors"; /root/reference/inference_codeformer.py:209-211).
static std::mutex g_status_mu;
static unsigned* g_status_words = nullptr;            // host pointer of a mapped, portable allocation: [64] words
static uint64_t g_status_bound = 0;                   // devices whose symbols are bound
static long long g_wait_limit_cycles = 4000000000LL;

int async_status_init(cudaStream_t) {
  int dev = 0;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_REQUIRE(dev >= 0 && dev < 64, "more than 64 CUDA devices are not supported");
  std::lock_guard<std::mutex> lk(g_status_mu);
  if (g_status_bound & (1ull << dev)) return 0;
  if (!g_status_words) {
    CFB_CUDA(cudaHostAlloc((void**)&g_status_words, 64 * sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(g_status_words, 0, 64 * sizeof(unsigned));
  }
  unsigned* dptr = nullptr;
  CFB_CUDA(cudaHostGetDevicePointer((void**)&dptr, g_status_words, 0));
  CFB_CHECK(tc_bind_status_word(dptr + dev, g_wait_limit_cycles));
  g_status_bound |= 1ull << dev;
  return 0;
}

int async_status_check(const char* where) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 0; }
  unsigned bits = 0;
  {
    std::lock_guard<std::mutex> lk(g_status_mu);
    if (!g_status_words || !(g_status_bound & (1ull << dev))) return 0;
    bits = __atomic_exchange_n(g_status_words + dev, 0u, __ATOMIC_ACQ_REL);
  }
  if (!bits) return 0;
  std::string msg = std::string(where) + ": a kernel of an earlier launch on this device reported";
  if (bits & CFB_STATUS_TIMEOUT) msg += " [barrier time-out: the tensor-core pipeline was aborted, results of that launch are invalid]";
  if (bits & CFB_STATUS_OVERFLOW) msg += " [fp16 operand overflow: an activation exceeded 65504 on the split-fp16 tensor-core path]";
  if (bits & CFB_STATUS_TIMEOUT) {
    if (cudaDeviceSynchronize() != cudaSuccess) cudaGetLastError();   // the aborted launch has drained; the context is healthy
    tc_clear_abort();
  }
  set_error(msg);
  return 1;
}

// ---- stream-ordered arena over the caller's workspace ---------------------------------------------
// All work of one forward is enqueued on one stream, so a block can be handed out again as soon as the
// host has *enqueued* its last reader.  First-fit with coalescing; `dry` mode only tracks the high-water mark.
class Arena {
 public:
  void reset(void* base, size_t cap, bool dry) {
    base_ = (char*)base; cap_ = cap; dry_ = dry; high_ = 0; free_.clear(); used_.clear();
    if (dry) { base_ = (char*)(uintptr_t)0x10000; cap_ = (size_t)1 << 46; }
    free_[0] = cap_;
  }
  void* alloc(size_t bytes) {
    bytes = (bytes + 1023) / 1024 * 1024;   // 1 KiB granularity keeps every tensor TMA/float4 aligned
    if (bytes == 0) bytes = 1024;
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= bytes) {
        const size_t off = it->first, sz = it->second;
        free_.erase(it);
        if (sz > bytes) free_[off + bytes] = sz - bytes;
        used_[off] = bytes;
        if (off + bytes > high_) high_ = off + bytes;
        return base_ + off;
      }
    }
    return nullptr;
  }
  void release(void* p) {
    if (!p) return;
    const size_t off = (size_t)((char*)p - base_);
    auto u = used_.find(off);
    if (u == used_.end()) return;
    size_t sz = u->second;
    used_.erase(u);
    auto nxt = free_.lower_bound(off);
    if (nxt != free_.end() && off + sz == nxt->first) { sz += nxt->second; nxt = free_.erase(nxt); }
    if (nxt != free_.begin()) {
      auto prv = std::prev(nxt);
      if (prv->first + prv->second == off) { prv->second += sz; return; }
    }
    free_[off] = sz;
  }
  size_t high() const { return high_; }
  bool dry() const { return dry_; }

 private:
  char* base_ = nullptr;
  size_t cap_ = 0, high_ = 0;
  bool dry_ = false;
  std::map<size_t, size_t> free_, used_;
};

struct Tensor {
  float* p = nullptr;
  int N = 0, H = 0, W = 0, C = 0;
  bool owned = true;  // false: caller memory or kept-alive tap
  float* gn_part = nullptr;   // GroupNorm partial sums emitted by the producing tensor-core conv (or null)
  int gn_slots = 0;           // partial slots per image
  void* planes = nullptr;     // fp16 hi/lo operand planes of this tensor emitted by the producing conv (raw values)
  const float* p2 = nullptr;  // channel concatenation held as two tensors: channels [0,C1) live in p, [C1,C) in p2 (Fuse_sft_block)
  int C1 = 0;
  int64_t numel() const { return (int64_t)N * H * W * C; }
};

struct ConvW {
  std::string name;
  int cin = 0, cout = 0, k = 0;
  bool has_bias = true;
  const float* src_w = nullptr;  // reference-layout source (row offset into a larger OIHW tensor allowed)
  const float* src_b = nullptr;
  float* w_f32 = nullptr;        // [taps][cin][cout]
  __half* w_hi = nullptr;        // [taps][cout][cin]
  __half* w_lo = nullptr;
  float* bias = nullptr;
  float* wscale = nullptr;       // 2 floats: [0] scratch |w|max, [1] 2^-k of the fp16 split
  bool is_up = false;            // conv of an Upsample block: also keep the 4-parity 2x2 weights for the tensor-core engine
  __half* u_hi = nullptr;        // [16][cout][cin]
  __half* u_lo = nullptr;
  float* uscale = nullptr;
};
struct NormW { std::string name; int c = 0; const float* gamma = nullptr; const float* beta = nullptr; };
struct ResW { NormW n1, n2; ConvW c1, c2, co; bool has_out = false; };
struct AttnW { NormW n; ConvW qkv, proj; float* consts = nullptr; /* device: [0] = C^-1/2, [1] = 1 */ };
struct FuseW { ResW enc; ConvW s0, s2, h0, h2; };
struct LayerW { NormW n1, n2; ConvW qk, v, o, l1, l2; };
struct Block { int kind; int cin, cout, res; ConvW conv; ResW res_w; AttnW attn; NormW norm; };
enum { B_CONV = 0, B_RES, B_ATTN, B_DOWN, B_UP, B_NORM };

}  // namespace cfb

using namespace cfb;

constexpr int GN_COUNTERS = 1 << 16;

struct cfb_net {
  cfb_config cfg;
  std::mutex mu;
  std::unordered_map<std::string, std::pair<const float*, int64_t>> raw;
  std::vector<Block> enc, gen;
  std::map<int, FuseW> fuse;          // keyed by feature size
  std::vector<LayerW> layers;
  ConvW feat_emb, idx_lin;
  ConvW vq_code;                      // codebook as a 1x1 'conv' (distance GEMM of VectorQuantizer.forward on tensor cores)
  NormW idx_norm;
  const float* position_emb = nullptr;
  const float* codebook = nullptr;    // points into slab copy
  float* slab = nullptr;
  size_t slab_bytes = 0;
  bool prepared = false;
  int64_t last_launches = 0;
  int sm_count = 148;
  bool tc_ok = false;                 // device is sm_100 => tcgen05 engine usable
  Arena arena;
  cudaStream_t st = nullptr;
  // small owned copies of norm params etc. live in the slab too
  std::vector<std::pair<const float**, std::pair<std::string, int64_t>>> vec_params;  // (dst, (name, numel))
  std::vector<ConvW*> convs;
  float* mha_consts = nullptr;        // device: [0] = head_dim^-1/2, [1] = 1
  unsigned* gn_counters = nullptr;    // ticket-counter ring of the split GroupNorm finalize (in the slab, zero between uses)
  int gn_ctr_pos = 0;
  int device = -1;                    // CUDA device the slab / prepared weights live on
  std::map<int, int64_t> ws_memo;     // batch -> cfb_workspace_bytes (16 host-side dry runs per miss)
  int engine = 0;                     // 0 auto (tcgen05 where the shape allows), 1 fp32 CUDA cores, 2 tcgen05 only
  std::map<std::string, std::pair<float*, int64_t>> captures;   // stage name -> (device dst, capacity in floats)
};

namespace cfb {

// ---- plan construction (mirrors the reference constructors) -----------------------------------------
static void mk_conv(cfb_net* n, ConvW& c, const std::string& name, int cin, int cout, int k, bool bias = true) {
  c.name = name; c.cin = cin; c.cout = cout; c.k = k; c.has_bias = bias;
  n->convs.push_back(&c);
}
static void mk_norm(cfb_net* n, NormW& w, const std::string& name, int c) {
  w.name = name; w.c = c;
  n->vec_params.push_back({&w.gamma, {name + ".weight", c}});
  n->vec_params.push_back({&w.beta, {name + ".bias", c}});
}
static void mk_res(cfb_net* n, ResW& r, const std::string& p, int cin, int cout) {
  mk_norm(n, r.n1, p + ".norm1", cin);
  mk_conv(n, r.c1, p + ".conv1", cin, cout, 3);
  mk_norm(n, r.n2, p + ".norm2", cout);
  mk_conv(n, r.c2, p + ".conv2", cout, cout, 3);
  r.has_out = cin != cout;
  if (r.has_out) mk_conv(n, r.co, p + ".conv_out", cin, cout, 1);
}
static bool in_list(const int32_t* l, int n, int v) {
  for (int i = 0; i < n; ++i) if (l[i] == v) return true;
  return false;
}

static void build_blocks(cfb_net* n, std::vector<Block>& blocks, const std::string& prefix,
                         const std::vector<std::array<int, 4>>& plan) {
  blocks.resize(plan.size());   // resize first: ConvW addresses are registered in n->convs
  for (size_t i = 0; i < plan.size(); ++i) {
    Block& b = blocks[i];
    b.kind = plan[i][0]; b.cin = plan[i][1]; b.cout = plan[i][2]; b.res = plan[i][3];
    const std::string p = prefix + ".blocks." + std::to_string(i);
    switch (b.kind) {
      case B_CONV: mk_conv(n, b.conv, p, b.cin, b.cout, 3); break;
      case B_RES: mk_res(n, b.res_w, p, b.cin, b.cout); break;
      case B_ATTN:
        mk_norm(n, b.attn.n, p + ".norm", b.cin);
        mk_conv(n, b.attn.qkv, p + ".qkv", b.cin, 3 * b.cin, 1);   // q,k,v fused along Cout (special-cased in prepare)
        mk_conv(n, b.attn.proj, p + ".proj_out", b.cin, b.cin, 1);
        break;
      case B_DOWN: case B_UP: mk_conv(n, b.conv, p + ".conv", b.cin, b.cout, 3); b.conv.is_up = (b.kind == B_UP); break;
      case B_NORM: mk_norm(n, b.norm, p, b.cin); break;
    }
  }
}

static int build_plan(cfb_net* n) {
  const cfb_config& c = n->cfg;
  CFB_REQUIRE(c.n_ch_mult >= 1 && c.n_ch_mult <= 8, "config: bad ch_mult length");
  CFB_REQUIRE(c.nf == 64, "config: only nf=64 is built (first/last conv kernels)");
  // Encoder.__init__  vqgan_arch.py:241-267
  std::vector<std::array<int, 4>> ep, gp;
  int curr = c.img_size;
  ep.push_back({B_CONV, 3, c.nf, curr});
  int cin = c.nf;
  for (int i = 0; i < c.n_ch_mult; ++i) {
    cin = c.nf * (i == 0 ? 1 : c.ch_mult[i - 1]);
    const int cout = c.nf * c.ch_mult[i];
    for (int r = 0; r < c.res_blocks; ++r) {
      ep.push_back({B_RES, cin, cout, curr});
      cin = cout;
      if (in_list(c.attn_res, c.n_attn_res, curr)) ep.push_back({B_ATTN, cin, cin, curr});
    }
    if (i != c.n_ch_mult - 1) { curr /= 2; ep.push_back({B_DOWN, cin, cin, curr}); }
  }
  ep.push_back({B_RES, cin, cin, curr});
  ep.push_back({B_ATTN, cin, cin, curr});
  ep.push_back({B_RES, cin, cin, curr});
  ep.push_back({B_NORM, cin, cin, curr});
  ep.push_back({B_CONV, cin, c.emb_dim, curr});
  // Generator.__init__  vqgan_arch.py:287-316
  cin = c.nf * c.ch_mult[c.n_ch_mult - 1];
  curr = c.img_size >> (c.n_ch_mult - 1);
  gp.push_back({B_CONV, c.emb_dim, cin, curr});
  gp.push_back({B_RES, cin, cin, curr});
  gp.push_back({B_ATTN, cin, cin, curr});
  gp.push_back({B_RES, cin, cin, curr});
  for (int i = c.n_ch_mult - 1; i >= 0; --i) {
    const int cout = c.nf * c.ch_mult[i];
    for (int r = 0; r < c.res_blocks; ++r) {
      gp.push_back({B_RES, cin, cout, curr});
      cin = cout;
      if (in_list(c.attn_res, c.n_attn_res, curr)) gp.push_back({B_ATTN, cin, cin, curr});
    }
    if (i != 0) { curr *= 2; gp.push_back({B_UP, cin, cin, curr}); }
  }
  gp.push_back({B_NORM, cin, cin, curr});
  gp.push_back({B_CONV, cin, 3, curr});
  build_blocks(n, n->enc, "encoder", ep);
  mk_conv(n, n->vq_code, "quantize.embedding", c.emb_dim, c.codebook_size, 1, false);
  build_blocks(n, n->gen, "generator", gp);
  if (c.kind == 1) {
    CFB_REQUIRE(c.img_size == 512 && c.emb_dim == 256, "config: CodeFormer is defined for 512x512 / emb 256");
    CFB_REQUIRE(c.dim_embd == 512 && c.dim_embd % c.n_head == 0 && c.dim_embd / c.n_head == 64,
                "config: only dim_embd=512 with 64-wide heads is built");
    mk_conv(n, n->feat_emb, "feat_emb", c.emb_dim, c.dim_embd, 1);
    n->layers.resize(c.n_layers);
    for (int l = 0; l < c.n_layers; ++l) {
      LayerW& L = n->layers[l];
      const std::string p = "ft_layers." + std::to_string(l);
      const int E = c.dim_embd;
      mk_conv(n, L.qk, p + ".self_attn.in_proj#qk", E, 2 * E, 1);
      mk_conv(n, L.v, p + ".self_attn.in_proj#v", E, E, 1);
      mk_conv(n, L.o, p + ".self_attn.out_proj", E, E, 1);
      mk_conv(n, L.l1, p + ".linear1", E, 2 * E, 1);
      mk_conv(n, L.l2, p + ".linear2", 2 * E, E, 1);
      mk_norm(n, L.n1, p + ".norm1", E);
      mk_norm(n, L.n2, p + ".norm2", E);
    }
    mk_norm(n, n->idx_norm, "idx_pred_layer.0", c.dim_embd);
    mk_conv(n, n->idx_lin, "idx_pred_layer.1", c.dim_embd, c.codebook_size, 1, false);
    static const int chan_of[6][2] = {{16, 512}, {32, 256}, {64, 256}, {128, 128}, {256, 128}, {512, 64}};
    for (int i = 0; i < c.n_connect; ++i) {
      int ch = 0;
      for (auto& e : chan_of) if (e[0] == c.connect[i]) ch = e[1];
      CFB_REQUIRE(ch != 0, "config: connect_list entries must be one of 16..512");
      FuseW& f = n->fuse[c.connect[i]];
      const std::string p = "fuse_convs_dict." + std::to_string(c.connect[i]);
      mk_res(n, f.enc, p + ".encode_enc", 2 * ch, ch);
      mk_conv(n, f.s0, p + ".scale.0", ch, ch, 3);
      mk_conv(n, f.s2, p + ".scale.2", ch, ch, 3);
      mk_conv(n, f.h0, p + ".shift.0", ch, ch, 3);
      mk_conv(n, f.h2, p + ".shift.2", ch, ch, 3);
    }
  }
  return 0;
}

// ---- weight preparation --------------------------------------------------------------------------------
static const float* find_param(cfb_net* n, const std::string& name, int64_t numel) {
  auto it = n->raw.find(name);
  if (it == n->raw.end()) { set_error("missing parameter '" + name + "'"); return nullptr; }
  if (it->second.second != numel) {
    set_error("parameter '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " +
              std::to_string(numel));
    return nullptr;
  }
  return it->second.first;
}

static int resolve_conv_sources(cfb_net* n, ConvW& c, float* qkv_scratch_w, float* qkv_scratch_b, cudaStream_t st) {
  const int64_t wn = (int64_t)c.cout * c.cin * c.k * c.k;
  const size_t hash = c.name.find('#');
  if (c.name.size() > 4 && c.name.compare(c.name.size() - 4, 4, ".qkv") == 0) {
    // AttnBlock q,k,v (vqgan_arch.py:173-193) stacked along Cout so one GEMM feeds the attention core
    const std::string p = c.name.substr(0, c.name.size() - 4);
    const int C = c.cin;
    const char* nm[3] = {".q", ".k", ".v"};
    for (int i = 0; i < 3; ++i) {
      const float* w = find_param(n, p + nm[i] + ".weight", (int64_t)C * C);
      const float* b = find_param(n, p + nm[i] + ".bias", C);
      if (!w || !b) return 1;
      CFB_CUDA(cudaMemcpyAsync(qkv_scratch_w + (int64_t)i * C * C, w, (size_t)C * C * 4, cudaMemcpyDeviceToDevice, st));
      CFB_CUDA(cudaMemcpyAsync(qkv_scratch_b + (int64_t)i * C, b, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    }
    c.src_w = qkv_scratch_w; c.src_b = qkv_scratch_b;
    return 0;
  }
  if (hash != std::string::npos) {
    // nn.MultiheadAttention in_proj_weight [3E,E] rows = [Wq;Wk;Wv]  (codeformer_arch.py:102)
    const std::string base = c.name.substr(0, hash);
    const std::string part = c.name.substr(hash + 1);
    const int E = c.cin;
    const float* w = find_param(n, base + "_weight", (int64_t)3 * E * E);
    const float* b = find_param(n, base + "_bias", (int64_t)3 * E);
    if (!w || !b) return 1;
    const int row0 = part == "qk" ? 0 : 2 * E;
    c.src_w = w + (int64_t)row0 * E; c.src_b = b + row0;
    return 0;
  }
  c.src_w = find_param(n, c.name + ".weight", wn);
  if (!c.src_w) return 1;
  if (c.has_bias) { c.src_b = find_param(n, c.name + ".bias", c.cout); if (!c.src_b) return 1; }
  return 0;
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static int prepare(cfb_net* n, cudaStream_t st) {
  // slab size
  size_t total = 0;
  for (ConvW* c : n->convs) {
    const size_t wn = (size_t)c->cout * c->cin * c->k * c->k;
    total += align256(wn * 4) + 2 * align256(wn * 2) + align256((size_t)c->cout * 4) + 256;
    if (c->is_up) total += 2 * align256((size_t)16 * c->cout * c->cin * 2) + 256;
  }
  for (auto& v : n->vec_params) total += align256((size_t)v.second.second * 4);
  total += align256((size_t)n->cfg.codebook_size * n->cfg.emb_dim * 4);
  if (n->cfg.kind == 1) total += align256((size_t)n->cfg.latent_size * n->cfg.dim_embd * 4);
  const size_t scratch = align256((size_t)3 * 512 * 512 * 4) + align256(3 * 512 * 4);
  total += scratch;
  total += 256 * (n->enc.size() + n->gen.size());      // per-AttnBlock device constants
  total += align256((size_t)GN_COUNTERS * sizeof(unsigned)) + 256;
  // the net lives on the device that is current at prepare time (net.to(other_gpu) -> a new prepare): slab, SM count and
  // engine availability all follow it
  int dev = 0, major = 0, sms = 148;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (n->slab && (n->device != dev || n->slab_bytes < total)) {
    if (n->device != dev && n->device >= 0) {
      cudaSetDevice(n->device);
      cudaFree(n->slab);
      cudaSetDevice(dev);
    } else {
      cudaFree(n->slab);
    }
    n->slab = nullptr; n->slab_bytes = 0;
  }
  if (!n->slab) {
    CFB_CUDA(cudaMalloc((void**)&n->slab, total));
    n->slab_bytes = total;
  }
  n->device = dev;
  n->sm_count = sms > 0 ? sms : 148;
  n->tc_ok = (major == 10);
  n->ws_memo.clear();
  CFB_CHECK(async_status_init(st));
  char* p = (char*)n->slab;
  auto take = [&](size_t bytes) { char* r = p; p += align256(bytes); return r; };
  n->mha_consts = (float*)take(256);
  {
    const float hc[2] = {n->cfg.kind == 1 ? 1.0f / sqrtf((float)(n->cfg.dim_embd / n->cfg.n_head)) : 1.0f, 1.0f};
    CFB_CUDA(cudaMemcpyAsync(n->mha_consts, hc, sizeof(hc), cudaMemcpyHostToDevice, st));
  }
  n->gn_counters = (unsigned*)take((size_t)GN_COUNTERS * sizeof(unsigned));
  n->gn_ctr_pos = 0;
  CFB_CUDA(cudaMemsetAsync(n->gn_counters, 0, (size_t)GN_COUNTERS * sizeof(unsigned), st));
  float* qkv_w = (float*)take((size_t)3 * 512 * 512 * 4);
  float* qkv_b = (float*)take(3 * 512 * 4);
  for (ConvW* c : n->convs) {
    if (c->name.size() > 4 && c->name.compare(c->name.size() - 4, 4, ".qkv") == 0)
      CFB_REQUIRE(c->cin <= 512, "AttnBlock wider than 512 channels is not built");
    CFB_CHECK(resolve_conv_sources(n, *c, qkv_w, qkv_b, st));
    const size_t wn = (size_t)c->cout * c->cin * c->k * c->k;
    c->w_f32 = (float*)take(wn * 4);
    c->w_hi = (__half*)take(wn * 2);
    c->w_lo = (__half*)take(wn * 2);
    c->bias = (float*)take((size_t)c->cout * 4);
    c->wscale = (float*)take(8);
    CFB_CHECK(relayout_oihw_to_tck(c->src_w, c->w_f32, c->cout, c->cin, c->k, st));
    CFB_CHECK(tc_split_weights(c->src_w, c->w_hi, c->w_lo, c->cout, c->cin, c->k, c->wscale, st));
    if (c->is_up) {
      c->u_hi = (__half*)take((size_t)16 * c->cout * c->cin * 2);
      c->u_lo = (__half*)take((size_t)16 * c->cout * c->cin * 2);
      c->uscale = (float*)take(8);
      CFB_CHECK(tc_split_weights_up4(c->src_w, c->u_hi, c->u_lo, c->cout, c->cin, c->uscale, st));
    }
    if (c->has_bias) CFB_CUDA(cudaMemcpyAsync(c->bias, c->src_b, (size_t)c->cout * 4, cudaMemcpyDeviceToDevice, st));
    else CFB_CUDA(cudaMemsetAsync(c->bias, 0, (size_t)c->cout * 4, st));
  }
  for (auto& v : n->vec_params) {
    const float* src = find_param(n, v.second.first, v.second.second);
    if (!src) return 1;
    float* dst = (float*)take((size_t)v.second.second * 4);
    CFB_CUDA(cudaMemcpyAsync(dst, src, (size_t)v.second.second * 4, cudaMemcpyDeviceToDevice, st));
    *v.first = dst;
  }
  {
    const int64_t ne = (int64_t)n->cfg.codebook_size * n->cfg.emb_dim;
    const float* src = find_param(n, "quantize.embedding.weight", ne);
    if (!src) return 1;
    float* dst = (float*)take((size_t)ne * 4);
    CFB_CUDA(cudaMemcpyAsync(dst, src, (size_t)ne * 4, cudaMemcpyDeviceToDevice, st));
    n->codebook = dst;
  }
  if (n->cfg.kind == 1) {
    const int64_t ne = (int64_t)n->cfg.latent_size * n->cfg.dim_embd;
    const float* src = find_param(n, "position_emb", ne);
    if (!src) return 1;
    float* dst = (float*)take((size_t)ne * 4);
    CFB_CUDA(cudaMemcpyAsync(dst, src, (size_t)ne * 4, cudaMemcpyDeviceToDevice, st));
    n->position_emb = dst;
  }
  for (std::vector<Block>* bl : {&n->enc, &n->gen})
    for (Block& b : *bl)
      if (b.kind == B_ATTN) {
        b.attn.consts = (float*)take(8);
        const float hc[2] = {1.0f / sqrtf((float)b.cin), 1.0f};
        CFB_CUDA(cudaMemcpyAsync(b.attn.consts, hc, sizeof(hc), cudaMemcpyHostToDevice, st));
      }
  // the qkv scratch is read by kernels enqueued above: the sources must stay valid until they ran
  CFB_CUDA(cudaStreamSynchronize(st));
  n->prepared = true;
  return 0;
}

// ---- forward building blocks ----------------------------------------------------------------------------
struct Fwd {
  cfb_net* n;
  cudaStream_t st;
  Arena& ar;
  bool dry;
  int engine;   // 0 auto, 1 f32, 2 tc
  // caller-side image plumbing fused into the first / last conv (cfb_codeformer_forward_u8): uint8 HWC BGR faces
  const unsigned char* in_u8 = nullptr;
  unsigned char* out_u8 = nullptr;

  int alloc(Tensor& t, int N, int H, int W, int C) {
    t.N = N; t.H = H; t.W = W; t.C = C; t.owned = true; t.gn_part = nullptr; t.gn_slots = 0; t.planes = nullptr;
    t.p2 = nullptr; t.C1 = 0;
    t.p = (float*)ar.alloc((size_t)t.numel() * 4);
    CFB_REQUIRE(t.p != nullptr, "workspace too small (use cfb_workspace_bytes)");
    return 0;
  }
  int alloc_raw(void** p, size_t bytes) {
    *p = ar.alloc(bytes);
    CFB_REQUIRE(*p != nullptr, "workspace too small (use cfb_workspace_bytes)");
    return 0;
  }
  void release(Tensor& t) {
    if (t.owned) {
      if (t.p) ar.release(t.p);
      if (t.gn_part) ar.release(t.gn_part);
      if (t.planes) ar.release(t.planes);
    }
    t.p = nullptr; t.gn_part = nullptr; t.planes = nullptr;
  }
  void release_raw(void* p) { ar.release(p); }

  // debug/parity hook: copy a stage's NHWC activation out (cfb_net_capture)
  int capture(const std::string& stage, const Tensor& t) {
    if (dry || n->captures.empty()) return 0;
    auto it = n->captures.find(stage);
    if (it == n->captures.end()) return 0;
    CFB_REQUIRE(it->second.second >= t.numel(), "capture buffer too small for stage " + stage);
    CFB_CUDA(cudaMemcpyAsync(it->second.first, t.p, (size_t)t.numel() * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  }

  struct ConvOpt {
    int mode = CONV_SAME;
    const float* in_scale = nullptr; const float* in_shift = nullptr; int in_act = IN_NONE;
    const float* residual = nullptr; int out_act = OUT_NONE;
    const float* sft_dec = nullptr; const float* sft_scale = nullptr; float sft_w = 0.f;
    float* out_ptr = nullptr;   // write into caller memory instead of the arena
    bool want_stats = false;    // consumer is a GroupNorm: let the tensor-core epilogue emit the partial sums
    bool want_planes = false;   // a following conv reads this output raw: emit its fp16 hi/lo operand planes too
    bool planes_only = false;   // ... and nothing reads the fp32 tensor: skip its store (tensor engine only)
  };

  int conv(const ConvW& w, const Tensor& in, Tensor& out, const ConvOpt& o) {
    CFB_REQUIRE(in.C == w.cin, "conv: channel mismatch for " + w.name);
    int Ho = in.H, Wo = in.W;
    if (o.mode == CONV_DOWN) { Ho = in.H / 2; Wo = in.W / 2; }
    if (o.mode == CONV_UP) { Ho = in.H * 2; Wo = in.W * 2; }
    ConvArgs a;
    a.in = in.p; a.N = in.N; a.H = in.H; a.W = in.W; a.Cin = in.C; a.Ho = Ho; a.Wo = Wo; a.Cout = w.cout;
    a.ksize = w.k; a.mode = o.mode; a.wgt_f32 = w.w_f32; a.wgt_hi = w.w_hi; a.wgt_lo = w.w_lo; a.wscale_inv = w.wscale + 1; a.bias = w.bias;
    a.in_scale = o.in_scale; a.in_shift = o.in_shift; a.in_act = o.in_act; a.residual = o.residual;
    a.out_act = o.out_act; a.sft_dec = o.sft_dec; a.sft_scale = o.sft_scale; a.sft_w = o.sft_w;
    bool use_tc = engine == 2 || (engine == 0 && n->tc_ok && tc_supported(a));
    const bool no_f32 = o.planes_only && o.want_planes && use_tc && !o.out_ptr;
    if (o.out_ptr) {
      out.p = o.out_ptr; out.N = in.N; out.H = Ho; out.W = Wo; out.C = w.cout; out.owned = false;
      out.gn_part = nullptr; out.gn_slots = 0; out.planes = nullptr; out.p2 = nullptr; out.C1 = 0;
    } else if (no_f32) {
      out.p = nullptr; out.N = in.N; out.H = Ho; out.W = Wo; out.C = w.cout; out.owned = true;
      out.gn_part = nullptr; out.gn_slots = 0; out.planes = nullptr; out.p2 = nullptr; out.C1 = 0;
    } else {
      CFB_CHECK(alloc(out, in.N, Ho, Wo, w.cout));
    }
    a.out = out.p;
    if (use_tc && o.mode == CONV_UP) { a.wgt_hi = w.u_hi; a.wgt_lo = w.u_lo; a.wscale_inv = w.uscale + 1; }
    if (use_tc) {
      CFB_REQUIRE(tc_supported(a), "conv: shape not supported by the tcgen05 engine: " + w.name);
      if (o.want_stats && !o.out_ptr && tc_can_emit_stats(a)) {
        out.gn_slots = tc_tiles_per_image(a) * 4;
        CFB_CHECK(alloc_raw((void**)&out.gn_part, (size_t)in.N * out.gn_slots * 64 * sizeof(float)));
        a.gn_part = out.gn_part;
      }
      if (o.want_planes && !o.out_ptr) {
        const size_t pb = (((size_t)in.N * Ho * Wo * w.cout * 2 + 1023) / 1024 * 1024) * 2;
        CFB_CHECK(alloc_raw(&out.planes, pb));
        a.out_planes = out.planes;
      }
      // Three ways the A operand reaches the tensor core:
      //  xf   GroupNorm-affine (+ SiLU) consumers on the halo + pair engine read the fp32 activation itself -- or the two
      //       halves of a channel concatenation -- and transform + split it inside the conv kernel (tc_can_xform);
      //  raw  the producer already emitted this tensor's fp16 hi/lo planes and the consumer takes it untransformed;
      //  prep everything else: a separate operand-preparation pass over the fp32 tensor.
      const bool plain = !o.in_scale && !o.in_shift && o.in_act == IN_NONE;
      a.halo1x1 = w.k == 1 && o.mode == CONV_SAME && o.in_scale && o.in_shift && in.p;     // AttnBlock q,k,v on GroupNorm(x)
      if (a.halo1x1 && !tc_can_xform(a)) a.halo1x1 = false;
      const bool xf = in.p && tc_can_xform(a) && ((o.in_scale && o.in_shift) || (plain && !in.planes));   // plain: in-kernel split only
      const bool raw = !xf && in.planes && plain;
      const bool reuse = raw || xf;
      void* scratch = raw ? in.planes : nullptr;
      if (!reuse) CFB_CHECK(alloc_raw(&scratch, tc_scratch_bytes(a)));
      CFB_REQUIRE(in.p != nullptr || raw, "conv: planes-only input without a planes consumer: " + w.name);
      CFB_REQUIRE(in.p2 == nullptr || reuse, "conv: a two-source tensor needs the fused operand transform or its planes: " + w.name);
      if (xf) { a.in2 = in.p2; a.Cin1 = in.C1; }
      a.skip_prep = reuse;
      a.xform = xf;
      if (!dry) CFB_CHECK(conv_tc(a, scratch, n->sm_count, st));
      if (!reuse) release_raw(scratch);
    } else {
      CFB_REQUIRE(in.p != nullptr && out.p != nullptr && in.p2 == nullptr, "conv: planes-only / two-source tensor reached the fp32 engine: " + w.name);
      if (!dry) CFB_CHECK(conv_f32(a, st));
    }
    return 0;
  }

  // GroupNorm(32, C, 1e-6) statistics -> scale/shift [N,C]
  int gn(const NormW& w, const Tensor& x, float** scale, float** shift) {
    CFB_REQUIRE(x.C == w.c, "norm: channel mismatch for " + w.name);
    CFB_REQUIRE(x.p || x.gn_part, "norm: planes-only tensor without partial sums for " + w.name);
    CFB_CHECK(alloc_raw((void**)scale, (size_t)x.N * x.C * 4));
    CFB_CHECK(alloc_raw((void**)shift, (size_t)x.N * x.C * 4));
    if (x.gn_part) {   // statistics already reduced per tile by the producing conv's epilogue
      void* scr = nullptr;
      const size_t sb = gn_final_scratch_bytes(x.N, x.gn_slots);
      if (sb) CFB_CHECK(alloc_raw(&scr, sb));
      if (!dry) {
        CFB_REQUIRE(x.N <= GN_COUNTERS / 2, "GroupNorm finalize: batch larger than the ticket-counter ring");
        // ticket counters: a ring in the net's slab, zero between uses (the kernel resets them); every call takes the next N
        // so that forwards in flight on other streams never share a counter
        if (n->gn_ctr_pos + x.N > GN_COUNTERS) n->gn_ctr_pos = 0;
        unsigned* ctr = n->gn_counters + n->gn_ctr_pos;
        n->gn_ctr_pos += x.N;
        CFB_CHECK(gn_coef_from_partials(x.gn_part, x.gn_slots, w.gamma, w.beta, *scale, *shift, x.N, x.H * x.W, x.C, 32, 1e-6f, scr, ctr, st));
      }
      if (scr) release_raw(scr);
      return 0;
    }
    void* ws = nullptr;
    CFB_CHECK(alloc_raw(&ws, gn_workspace_bytes(x.N, x.H * x.W, x.C)));
    if (!dry) CFB_CHECK(gn_coef(x.p, w.gamma, w.beta, *scale, *shift, x.N, x.H * x.W, x.C, 32, 1e-6f, ws, st));
    release_raw(ws);
    return 0;
  }

  // ResBlock.forward  vqgan_arch.py:153-164
  int resblock(const ResW& r, const Tensor& x, Tensor& y, bool out_planes = false) {
    float *s1, *h1, *s2, *h2;
    CFB_CHECK(gn(r.n1, x, &s1, &h1));
    Tensor h;
    ConvOpt o1; o1.in_scale = s1; o1.in_shift = h1; o1.in_act = IN_SILU; o1.want_stats = true;
    CFB_CHECK(conv(r.c1, x, h, o1));      // conv2 reads h as fp32 (fused operand transform or prep pass): no planes of h
    release_raw(s1); release_raw(h1);
    CFB_CHECK(gn(r.n2, h, &s2, &h2));
    Tensor skip = x; skip.owned = false;
    if (r.has_out) { ConvOpt oo; CFB_CHECK(conv(r.co, x, skip, oo)); }
    ConvOpt o2; o2.in_scale = s2; o2.in_shift = h2; o2.in_act = IN_SILU; o2.residual = skip.p; o2.want_stats = true;
    o2.want_planes = out_planes;
    CFB_CHECK(conv(r.c2, h, y, o2));
    release_raw(s2); release_raw(h2);
    release(h);
    if (r.has_out) release(skip);
    return 0;
  }

  // AttnBlock.forward  vqgan_arch.py:202-226
  int attnblock(const AttnW& w, const Tensor& x, Tensor& y, bool out_planes = false) {
    CFB_REQUIRE(x.H * x.W == 256, "AttnBlock: built for the 16x16 latent");
    float *s, *h;
    CFB_CHECK(gn(w.n, x, &s, &h));
    const int C = x.C;
    const bool tc_attn = engine != 1 && n->tc_ok && x.H == 16 && x.W == 16 && C % 128 == 0;
    Tensor qkv;
    ConvOpt o; o.in_scale = s; o.in_shift = h; o.want_planes = tc_attn;
    CFB_CHECK(conv(w.qkv, x, qkv, o));
    release_raw(s); release_raw(h);
    Tensor a;
    CFB_CHECK(alloc(a, x.N, x.H, x.W, x.C));
    if (tc_attn && qkv.planes) {
      // attention core on the tcgen05 engine: scores = q k^T C^-1/2 and out = P v as per-image GEMMs on operand planes
      const int64_t T = (int64_t)x.N * 256;
      float* scores = nullptr;
      void *pp = nullptr, *vt = nullptr;
      CFB_CHECK(alloc_raw((void**)&scores, (size_t)T * 256 * 4));
      CFB_CHECK(alloc_raw(&pp, 2 * (((size_t)T * 256 * 2 + 1023) / 1024 * 1024)));
      CFB_CHECK(alloc_raw(&vt, 2 * (((size_t)x.N * C * 256 * 2 + 1023) / 1024 * 1024)));
      CFB_CHECK(alloc_raw(&a.planes, 2 * (((size_t)T * C * 2 + 1023) / 1024 * 1024)));
      if (!dry) {
        BmmArgs g1;
        g1.a_planes = qkv.planes; g1.a_pitch = 3 * C; g1.a_c0 = 0;
        g1.b_planes = qkv.planes; g1.b_pitch = 3 * C; g1.b_c0 = C; g1.b_rows = 256;
        g1.N = x.N; g1.K = C; g1.Cout = 256; g1.scale_dev = w.consts; g1.out = scores;
        CFB_CHECK(bmm_tc(g1, n->sm_count, st));
        CFB_CHECK(softmax256_planes(scores, pp, T, st));
        CFB_CHECK(transpose_planes(qkv.planes, x.N, 3 * C, 2 * C, C, vt, st));
        BmmArgs g2;
        g2.a_planes = pp; g2.a_pitch = 256; g2.a_c0 = 0;
        g2.b_planes = vt; g2.b_pitch = 256; g2.b_c0 = 0; g2.b_rows = C;
        g2.N = x.N; g2.K = 256; g2.Cout = C; g2.scale_dev = w.consts + 1; g2.out = a.p; g2.out_planes = a.planes;
        CFB_CHECK(bmm_tc(g2, n->sm_count, st));
      }
      release_raw(scores); release_raw(pp); release_raw(vt);
    } else {
      if (!dry)
        CFB_CHECK(attention(qkv.p, qkv.p + C, qkv.p + 2 * C, a.p, x.N, 256, 1, C, 3 * C, 3 * C, 3 * C, C,
                            1.0f / sqrtf((float)C), st));
    }
    release(qkv);
    ConvOpt op; op.residual = x.p; op.want_stats = true; op.want_planes = out_planes;
    CFB_CHECK(conv(w.proj, a, y, op));
    release(a);
    return 0;
  }

  // Fuse_sft_block.forward  codeformer_arch.py:151-157
  int fuse(const FuseW& f, const Tensor& enc_feat, const Tensor& dec, float wgt, Tensor& y) {
    Tensor cat;
    const bool stats_from_parts = enc_feat.gn_part && dec.gn_part && enc_feat.gn_slots == dec.gn_slots && enc_feat.C == dec.C;
    bool two_src = false;
    if (stats_from_parts && (engine == 2 || (engine == 0 && n->tc_ok))) {
      ConvArgs a1;      // conv1 of the fused ResBlock: does it run the in-kernel operand transform?
      a1.N = dec.N; a1.H = dec.H; a1.W = dec.W; a1.Cin = f.enc.c1.cin; a1.Ho = dec.H; a1.Wo = dec.W; a1.Cout = f.enc.c1.cout;
      a1.ksize = f.enc.c1.k; a1.mode = CONV_SAME;
      two_src = f.enc.has_out && tc_can_xform(a1) && enc_feat.C % 64 == 0 && dec.C % 64 == 0;
    }
    void* cat_planes = nullptr;
    if (two_src) {
      // torch.cat([enc_feat, dec]) (codeformer_arch.py:152) is never materialised in fp32: conv1 of the fused ResBlock reads the
      // two tensors through two tensor maps (fused operand transform), GroupNorm statistics come from the sources' partial
      // sums, and only the raw 1x1 conv_out needs the concatenation -- as fp16 hi/lo operand planes
      cat.p = enc_feat.p; cat.p2 = dec.p; cat.C1 = enc_feat.C;
      cat.N = dec.N; cat.H = dec.H; cat.W = dec.W; cat.C = enc_feat.C + dec.C; cat.owned = false;
      cat.gn_part = nullptr; cat.gn_slots = 0; cat.planes = nullptr;
      CFB_CHECK(alloc_raw(&cat_planes, 2 * (((size_t)cat.numel() * 2 + 1023) / 1024 * 1024)));
      cat.planes = cat_planes;
      if (!dry) CFB_CHECK(concat_planes(enc_feat.p, dec.p, cat.planes, (int64_t)dec.N * dec.H * dec.W, enc_feat.C, dec.C, st));
    } else {
      CFB_CHECK(alloc(cat, dec.N, dec.H, dec.W, enc_feat.C + dec.C));
      if (!dry) CFB_CHECK(concat_channels(enc_feat.p, dec.p, cat.p, (int64_t)dec.N * dec.H * dec.W, enc_feat.C, dec.C, st));
    }
    float* cat_part = nullptr;
    if (stats_from_parts) {
      // GroupNorm statistics of the concatenation follow from the two sources' partial sums (no extra pass)
      cat.gn_slots = dec.gn_slots;
      CFB_CHECK(alloc_raw((void**)&cat_part, (size_t)dec.N * cat.gn_slots * 64 * sizeof(float)));
      cat.gn_part = cat_part;
      if (!dry) CFB_CHECK(gn_cat_partials(enc_feat.gn_part, dec.gn_part, cat.gn_part, (int64_t)dec.N * cat.gn_slots, st));
    }
    Tensor e;
    CFB_CHECK(resblock(f.enc, cat, e, true));      // scale.0 / shift.0 both read `e` raw: one set of planes, no prep
    if (two_src) { release_raw(cat_planes); if (cat_part) release_raw(cat_part); cat.planes = nullptr; cat.gn_part = nullptr; }
    else release(cat);
    Tensor s0, sc, h0;
    ConvOpt ol; ol.out_act = OUT_LRELU; ol.want_planes = true;      // s0 / h0 feed scale.2 / shift.2 raw
    CFB_CHECK(conv(f.s0, e, s0, ol));
    ConvOpt on;
    CFB_CHECK(conv(f.s2, s0, sc, on));
    release(s0);
    CFB_CHECK(conv(f.h0, e, h0, ol));
    release(e);
    ConvOpt of; of.sft_dec = dec.p; of.sft_scale = sc.p; of.sft_w = wgt; of.want_stats = true; of.want_planes = true;
    CFB_CHECK(conv(f.h2, h0, y, of));
    release(h0); release(sc);
    return 0;
  }

  // conv1 of a fused ResBlock at this shape: fused operand transform available?
  bool xf_ok(int N, int H, int W, const ConvW& w) const {
    if (!(engine == 2 || (engine == 0 && n->tc_ok))) return false;
    ConvArgs a;
    a.N = N; a.H = H; a.W = W; a.Ho = H; a.Wo = W; a.Cin = w.cin; a.Cout = w.cout; a.ksize = w.k; a.mode = CONV_SAME;
    return tc_can_xform(a);
  }
  // Does a consumer of block i's output read it RAW through the tensor engine (then the producer also emits its fp16 hi/lo
  // operand planes)?  Down/Upsample convs and a ResBlock's 1x1 conv_out do.  GroupNorm (+SiLU) consumers -- a ResBlock's
  // conv1, a norm -> conv pair -- read the fp32 tensor itself (fused operand transform or prep pass).
  bool next_takes_planes(const std::vector<Block>& bl, size_t i) const {
    if (!(engine == 2 || (engine == 0 && n->tc_ok))) return false;
    if (i + 1 >= bl.size()) return false;
    const Block& nb = bl[i + 1];
    if (nb.kind == B_DOWN || nb.kind == B_UP) return true;
    if (nb.kind == B_RES) return nb.res_w.has_out;
    return false;
  }

  // Encoder.forward (+ the taps of codeformer_arch.py:226-230).  x_nchw is the caller's image.
  int encoder(const float* x_nchw, int B, Tensor& z, std::map<int, Tensor>* taps, const std::vector<int>& tap_blocks) {
    const cfb_config& c = n->cfg;
    Tensor x;
    CFB_CHECK(alloc(x, B, c.img_size, c.img_size, c.nf));
    // on the tensor-core path every GroupNorm takes its statistics from partial sums of the producing kernel's epilogue: the
    // first conv emits them too (no pass over the 64-channel full-resolution tensor for the first ResBlock's norm1)
    const int64_t hw = (int64_t)c.img_size * c.img_size;
    if (engine != 1 && n->tc_ok && c.nf == 64 && hw % 256 == 0 && B <= GN_COUNTERS / 2) {
      x.gn_slots = (int)(hw / 32);
      CFB_CHECK(alloc_raw((void**)&x.gn_part, (size_t)B * x.gn_slots * 64 * sizeof(float)));
    }
    if (!dry) {
      if (in_u8) CFB_CHECK(conv_first_u8(in_u8, n->enc[0].conv.w_f32, n->enc[0].conv.bias, x.p, B, c.img_size, c.img_size, c.nf, st, x.gn_part));
      else CFB_CHECK(conv_first(x_nchw, n->enc[0].conv.w_f32, n->enc[0].conv.bias, x.p, B, c.img_size, c.img_size, c.nf, st, x.gn_part));
    }
    CFB_CHECK(capture("enc.0", x));
    float *ps = nullptr, *ph = nullptr;   // pending GroupNorm of a 'norm' block
    for (size_t i = 1; i < n->enc.size(); ++i) {
      const Block& b = n->enc[i];
      const bool pl = next_takes_planes(n->enc, i);
      Tensor y;
      switch (b.kind) {
        case B_RES: CFB_CHECK(resblock(b.res_w, x, y, pl)); break;
        case B_ATTN: CFB_CHECK(attnblock(b.attn, x, y, pl)); break;
        case B_DOWN: { ConvOpt o; o.mode = CONV_DOWN; o.want_stats = true; o.want_planes = pl; CFB_CHECK(conv(b.conv, x, y, o)); break; }
        case B_NORM: CFB_CHECK(gn(b.norm, x, &ps, &ph)); continue;   // consumed by the next conv
        case B_CONV: {
          ConvOpt o; o.in_scale = ps; o.in_shift = ph;
          o.want_planes = (i + 1 == n->enc.size()) && n->cfg.kind == 1;   // lq_feat feeds feat_emb raw
          CFB_CHECK(conv(b.conv, x, y, o));
          if (ps) { release_raw(ps); release_raw(ph); ps = ph = nullptr; }
          break;
        }
        default: CFB_REQUIRE(false, "encoder: unexpected block kind");
      }
      release(x);
      x = y;
      CFB_CHECK(capture("enc." + std::to_string(i), x));
      for (int tb : tap_blocks)
        if ((int)i == tb && taps) { x.owned = false; (*taps)[x.W] = x; (*taps)[x.W].owned = true; }
    }
    z = x;
    return 0;
  }

  // Generator.forward with the SFT fusion of codeformer_arch.py:272-277; writes NCHW into out_nchw
  int generator(Tensor x, float* out_nchw, std::map<int, Tensor>* taps, const std::vector<int>& fuse_blocks, float w) {
    float *ps = nullptr, *ph = nullptr;
    for (size_t i = 0; i < n->gen.size(); ++i) {
      const Block& b = n->gen[i];
      bool pl = next_takes_planes(n->gen, i);
      if (taps && w > 0.f)
        for (int fb : fuse_blocks)
          if ((int)i == fb) pl = false;      // consumed by the fusion (concat + SFT read fp32); the fused output emits its own
      Tensor y;
      switch (b.kind) {
        case B_RES: CFB_CHECK(resblock(b.res_w, x, y, pl)); break;
        case B_ATTN: CFB_CHECK(attnblock(b.attn, x, y, pl)); break;
        case B_UP: { ConvOpt o; o.mode = CONV_UP; o.want_stats = true; o.want_planes = pl; CFB_CHECK(conv(b.conv, x, y, o)); break; }
        case B_NORM: CFB_CHECK(gn(b.norm, x, &ps, &ph)); continue;
        case B_CONV:
          if (i + 1 == n->gen.size()) {
            if (!dry) {
              if (out_u8) CFB_CHECK(conv_last_u8(x.p, ps, ph, b.conv.w_f32, b.conv.bias, out_u8, x.N, x.H, x.W, x.C, st));
              else CFB_CHECK(conv_last(x.p, ps, ph, b.conv.w_f32, b.conv.bias, out_nchw, x.N, x.H, x.W, x.C, st));
            }
            if (ps) { release_raw(ps); release_raw(ph); ps = ph = nullptr; }
            release(x);
            return 0;
          } else {
            ConvOpt o; o.in_scale = ps; o.in_shift = ph; o.want_stats = true; o.want_planes = pl;
            CFB_CHECK(conv(b.conv, x, y, o));
            if (ps) { release_raw(ps); release_raw(ph); ps = ph = nullptr; }
          }
          break;
        default: CFB_REQUIRE(false, "generator: unexpected block kind");
      }
      release(x);
      x = y;
      CFB_CHECK(capture("gen." + std::to_string(i), x));
      if (taps && w > 0.f)
        for (int fb : fuse_blocks)
          if ((int)i == fb) {
            auto it = taps->find(x.W);
            CFB_REQUIRE(it != taps->end(), "fusion: encoder feature missing");
            auto fw = n->fuse.find(x.W);
            CFB_REQUIRE(fw != n->fuse.end(), "fusion: no Fuse_sft_block for this size");
            Tensor fz;
            CFB_CHECK(fuse(fw->second, it->second, x, w, fz));
            release(x);
            release(it->second);
            x = fz;
            CFB_CHECK(capture("fuse." + std::to_string(x.W), x));
          }
    }
    CFB_REQUIRE(false, "generator: plan does not end with a conv");
    return 1;
  }

  // TransformerSALayer.forward x9 + idx_pred_layer  codeformer_arch.py:235-245
  int transformer(const Tensor& lq, float* logits_out) {
    const cfb_config& c = n->cfg;
    const int B = lq.N, S = lq.H * lq.W, E = c.dim_embd, T = B * S;
    CFB_REQUIRE(S == c.latent_size, "transformer: token count != latent_size");
    Tensor tok = lq; tok.owned = false;   // [B,16,16,256] == tokens [T,256]
    Tensor x;
    { ConvOpt o; CFB_CHECK(conv(n->feat_emb, tok, x, o)); }
    // On the tensor engine every LayerNorm / attention / GELU output of a layer is only ever a GEMM operand: the producing
    // kernel writes it straight as fp16 hi/lo operand planes (no fp32 copy, no operand-preparation pass).
    const bool tcp = engine != 1 && n->tc_ok;
    const size_t plE = 2 * (((size_t)T * E * 2 + 1023) / 1024 * 1024);
    auto planes_tensor = [&](Tensor& t, int C, size_t bytes) -> int {
      t.p = nullptr; t.N = B; t.H = lq.H; t.W = lq.W; t.C = C; t.owned = true; t.gn_part = nullptr; t.gn_slots = 0;
      t.p2 = nullptr; t.C1 = 0; t.planes = nullptr;
      return alloc_raw(&t.planes, bytes);
    };
    for (const LayerW& L : n->layers) {
      Tensor t2, qkin, qk, v, a, x2, hdn, x3;
      if (tcp) {
        CFB_CHECK(planes_tensor(t2, E, plE));
        CFB_CHECK(planes_tensor(qkin, E, plE));
        if (!dry) CFB_CHECK(layer_norm_planes(x.p, L.n1.gamma, L.n1.beta, t2.planes, qkin.planes, n->position_emb, S, T, E, st));
      } else {
        CFB_CHECK(alloc(t2, B, lq.H, lq.W, E));
        CFB_CHECK(alloc(qkin, B, lq.H, lq.W, E));
        if (!dry) CFB_CHECK(layer_norm(x.p, L.n1.gamma, L.n1.beta, t2.p, qkin.p, n->position_emb, S, T, E, st));
      }
      { ConvOpt o; o.want_planes = tcp; o.planes_only = tcp; CFB_CHECK(conv(L.qk, qkin, qk, o)); }
      { ConvOpt o; o.want_planes = tcp; o.planes_only = tcp; CFB_CHECK(conv(L.v, t2, v, o)); }
      release(qkin); release(t2);
      if (tcp) {
        // nn.MultiheadAttention core (codeformer_arch.py:126) on the tcgen05 engine: per (image, head) GEMMs on operand planes --
        // scores = (q_h k_h^T) * d^-1/2 (K = 64, the power-of-two scale commutes exactly with the product), softmax -> planes
        // of the probabilities, out_h = P v_h written as the operand planes of out_proj (no fp32 copy of anything)
        const int Hh = c.n_head, dh = E / c.n_head;
        CFB_CHECK(planes_tensor(a, E, plE));
        float* scores = nullptr;
        void *pp = nullptr, *vt = nullptr;
        const int64_t rows = (int64_t)B * Hh * S;
        CFB_CHECK(alloc_raw((void**)&scores, (size_t)rows * S * 4));
        CFB_CHECK(alloc_raw(&pp, 2 * (((size_t)rows * S * 2 + 1023) / 1024 * 1024)));
        CFB_CHECK(alloc_raw(&vt, 2 * (((size_t)B * E * S * 2 + 1023) / 1024 * 1024)));
        if (!dry) {
          BmmArgs g1;
          g1.a_planes = qk.planes; g1.a_pitch = 2 * E; g1.a_c0 = 0; g1.a_c_head = dh;
          g1.b_planes = qk.planes; g1.b_pitch = 2 * E; g1.b_c0 = E; g1.b_c_head = dh; g1.b_rows = S;
          g1.N = B; g1.heads = Hh; g1.K = dh; g1.Cout = S; g1.scale_dev = n->mha_consts; g1.out = scores; g1.out_per_head = true;
          CFB_CHECK(bmm_tc(g1, n->sm_count, st));
          CFB_CHECK(softmax256_planes(scores, pp, rows, st));
          CFB_CHECK(transpose_planes(v.planes, B, E, 0, E, vt, st));
          BmmArgs g2;
          g2.a_planes = pp; g2.a_pitch = S; g2.a_c0 = 0; g2.a_img_per_head = true;
          g2.b_planes = vt; g2.b_pitch = S; g2.b_c0 = 0; g2.b_rows = E; g2.b_r_head = dh;
          g2.N = B; g2.heads = Hh; g2.K = S; g2.Cout = dh; g2.scale_dev = n->mha_consts + 1; g2.out = nullptr; g2.out_planes = a.planes;
          g2.out_per_head = false; g2.o_c_head = dh;
          CFB_CHECK(bmm_tc(g2, n->sm_count, st));
        }
        release_raw(scores); release_raw(pp); release_raw(vt);
      } else {
        CFB_CHECK(alloc(a, B, lq.H, lq.W, E));
        if (!dry)
          CFB_CHECK(attention(qk.p, qk.p + E, v.p, a.p, B, S, c.n_head, E / c.n_head, 2 * E, 2 * E, E, E,
                              sqrtf(1.0f / (float)(E / c.n_head)), st, nullptr));
      }
      release(qk); release(v);
      { ConvOpt o; o.residual = x.p; CFB_CHECK(conv(L.o, a, x2, o)); }
      release(a); release(x);
      if (tcp) {
        CFB_CHECK(planes_tensor(t2, E, plE));
        if (!dry) CFB_CHECK(layer_norm_planes(x2.p, L.n2.gamma, L.n2.beta, t2.planes, nullptr, nullptr, 0, T, E, st));
      } else {
        CFB_CHECK(alloc(t2, B, lq.H, lq.W, E));
        if (!dry) CFB_CHECK(layer_norm(x2.p, L.n2.gamma, L.n2.beta, t2.p, nullptr, nullptr, 0, T, E, st));
      }
      { ConvOpt o; o.out_act = OUT_GELU; o.want_planes = tcp; o.planes_only = tcp; CFB_CHECK(conv(L.l1, t2, hdn, o)); }
      release(t2);
      { ConvOpt o; o.residual = x2.p; CFB_CHECK(conv(L.l2, hdn, x3, o)); }
      release(hdn); release(x2);
      x = x3;
      CFB_CHECK(capture("ft." + std::to_string((int)(&L - &n->layers[0])), x));
    }
    Tensor t2, lg;
    if (tcp) {
      CFB_CHECK(planes_tensor(t2, E, plE));
      if (!dry) CFB_CHECK(layer_norm_planes(x.p, n->idx_norm.gamma, n->idx_norm.beta, t2.planes, nullptr, nullptr, 0, T, E, st));
    } else {
      CFB_CHECK(alloc(t2, B, lq.H, lq.W, E));
      if (!dry) CFB_CHECK(layer_norm(x.p, n->idx_norm.gamma, n->idx_norm.beta, t2.p, nullptr, nullptr, 0, T, E, st));
    }
    release(x);
    { ConvOpt o; o.out_ptr = logits_out; CFB_CHECK(conv(n->idx_lin, t2, lg, o)); }
    release(t2);
    return 0;
  }
};

// The prepared weights live on ONE device; a forward issued while another device is current would launch kernels there
// on foreign memory.  Also the place where an asynchronous failure of an earlier launch is reported (never silently lost).
static int check_device(cfb_net* n) {
  int dev = -1;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_REQUIRE(dev == n->device, "the net was prepared on CUDA device " + std::to_string(n->device) + " but device " +
                                    std::to_string(dev) + " is current (call net.to(device) / cfb_net_prepare again)");
  return async_status_check("forward");
}

static std::vector<int> tap_blocks_of(const cfb_config& c, bool encoder) {
  // fuse_encoder_block / fuse_generator_block  codeformer_arch.py:204-206
  static const int enc_of[6][2] = {{512, 2}, {256, 5}, {128, 8}, {64, 11}, {32, 14}, {16, 18}};
  static const int gen_of[6][2] = {{16, 6}, {32, 9}, {64, 12}, {128, 15}, {256, 18}, {512, 21}};
  std::vector<int> r;
  for (int i = 0; i < c.n_connect; ++i)
    for (auto& e : (encoder ? enc_of : gen_of))
      if (e[0] == c.connect[i]) r.push_back(e[1]);
  return r;
}

static int codeformer_forward_impl(cfb_net* n, const float* x, float* out, float* logits, float* lq_feat,
                                   int64_t* top_idx, int B, float w, int adain, int code_only, void* ws, int64_t ws_bytes,
                                   cudaStream_t st, bool dry, const unsigned char* x_u8 = nullptr,
                                   unsigned char* out_u8 = nullptr) {
  CFB_REQUIRE(n->cfg.kind == 1, "net was created as VQAutoEncoder");
  CFB_REQUIRE(dry || n->prepared, "cfb_net_prepare has not been called");
  if (!dry) CFB_CHECK(check_device(n));
  CFB_REQUIRE(B >= 0, "negative batch");
  if (B == 0) return 0;
  n->arena.reset(ws, (size_t)ws_bytes, dry);
  Fwd f{n, st, n->arena, dry, n->engine};
  f.in_u8 = x_u8; f.out_u8 = out_u8;
  const cfb_config& c = n->cfg;
  std::map<int, Tensor> taps;
  Tensor lq;
  const bool want_taps = (w > 0.f) && !code_only;   // codeformer_arch.py:276 -- features are only consumed when w>0
  CFB_CHECK(f.encoder(x, B, lq, want_taps ? &taps : nullptr, tap_blocks_of(c, true)));
  const int T = B * lq.H * lq.W;
  float* logits_buf = logits;
  if (!logits_buf) CFB_CHECK(f.alloc_raw((void**)&logits_buf, (size_t)T * c.codebook_size * 4));
  CFB_CHECK(f.transformer(lq, logits_buf));
  if (lq_feat && !dry) CFB_CHECK(nhwc_to_nchw(lq.p, lq_feat, B, lq.C, lq.H * lq.W, st));
  if (code_only) {                                     // codeformer_arch.py:247-249
    if (top_idx && !dry) CFB_CHECK(argmax_gather(logits_buf, n->codebook, top_idx, nullptr, T, c.codebook_size, c.emb_dim, st));
    return 0;
  }
  CFB_REQUIRE(out != nullptr || out_u8 != nullptr, "out must not be NULL unless code_only");
  // softmax -> topk(1) -> get_codebook_feat  (:257-259)
  Tensor quant;
  CFB_CHECK(f.alloc(quant, B, lq.H, lq.W, c.emb_dim));
  if (!dry) CFB_CHECK(argmax_gather(logits_buf, n->codebook, top_idx, quant.p, T, c.codebook_size, c.emb_dim, st));
  if (adain) {                                         // :265-266
    Tensor q2;
    CFB_CHECK(f.alloc(q2, B, lq.H, lq.W, c.emb_dim));
    if (!dry) CFB_CHECK(adain_nhwc(quant.p, lq.p, q2.p, B, lq.H * lq.W, c.emb_dim, st));
    f.release(quant);
    quant = q2;
  }
  CFB_CHECK(f.capture("quant", quant));
  f.release(lq);
  CFB_CHECK(f.generator(quant, out, want_taps ? &taps : nullptr, tap_blocks_of(c, false), w));
  return 0;
}

static int vqae_forward_impl(cfb_net* n, const float* x, float* out, int64_t* idx, float* stats, float* onehot, int B,
                             void* ws, int64_t ws_bytes, cudaStream_t st, bool dry) {
  CFB_REQUIRE(dry || n->prepared, "cfb_net_prepare has not been called");
  if (B == 0) return 0;
  if (!dry) CFB_CHECK(check_device(n));
  n->arena.reset(ws, (size_t)ws_bytes, dry);
  Fwd f{n, st, n->arena, dry, n->engine};
  const cfb_config& c = n->cfg;
  Tensor z;
  CFB_CHECK(f.encoder(x, B, z, nullptr, {}));
  const int T = B * z.H * z.W;
  Tensor zq;
  CFB_CHECK(f.alloc(zq, B, z.H, z.W, z.C));
  int64_t* idx_buf = idx;
  float* stats_buf = stats;
  if (!idx_buf) CFB_CHECK(f.alloc_raw((void**)&idx_buf, (size_t)T * 8));
  if (!stats_buf) CFB_CHECK(f.alloc_raw((void**)&stats_buf, 16));
  bool tc_vq = false;
  {
    ConvArgs probe;
    probe.N = B; probe.H = z.H; probe.W = z.W; probe.Cin = z.C; probe.Ho = z.H; probe.Wo = z.W; probe.Cout = c.codebook_size;
    probe.ksize = 1; probe.mode = CONV_SAME;
    tc_vq = n->engine != 1 && n->tc_ok && tc_supported(probe);
  }
  if (tc_vq) {
    // distance GEMM z.E^T on the tcgen05 engine, then the warp-shuffle argmin over the dot products
    Tensor dots;
    Fwd::ConvOpt o;
    CFB_CHECK(f.conv(n->vq_code, z, dots, o));
    void* vws = nullptr;
    CFB_CHECK(f.alloc_raw(&vws, vq_select_workspace_bytes(T, c.codebook_size)));
    if (!dry)
      CFB_CHECK(vq_select_from_dots(z.p, n->codebook, dots.p, T, z.C, c.codebook_size, c.beta, idx_buf, zq.p, stats_buf, onehot, vws, st));
    f.release(dots);
  } else {
    void* vws = nullptr;
    CFB_CHECK(f.alloc_raw(&vws, vq_workspace_bytes(T, z.C, c.codebook_size)));
    if (!dry) CFB_CHECK(vq_nearest(z.p, n->codebook, T, z.C, c.codebook_size, c.beta, idx_buf, zq.p, stats_buf, onehot, vws, st));
  }
  f.release(z);
  CFB_CHECK(f.generator(zq, out, nullptr, {}, 0.f));
  return 0;
}

}  // namespace cfb

// =========================================================================================================
// RRDBNet: the background / face upsampler behind RealESRGANer.enhance
//   /root/reference/basicsr/archs/rrdbnet_arch.py:9-120, call sites /root/reference/basicsr/utils/realesrgan_utils.py:100-175
// 23 RRDBs of three ResidualDenseBlocks: every 3x3 conv runs on the tcgen05 engine in its generalised fused-transform form
// (fp32 activation read in place, fp16 hi/lo split inside the kernel, any H x W, zero padding by TMA out-of-bounds fill).
// The dense concatenations torch.cat((x, x1, ..)) never exist: one NHWC buffer of 192 channels per block holds
// [x | x1 | x2 | x3 | x4]; conv_k reads its 64-aligned channel window (weights are zero beyond the real Cin) and writes its
// 32 growth channels at their offset; conv5 writes `x5*0.2 + x` (0.2 folded into the weight scale and bias) into the next
// block's buffer, and the third block of an RRDB also applies `*0.2 + x_rrdb` in the same epilogue.
// =========================================================================================================
namespace cfb {
struct GenConv {
  std::string name;
  int cin = 0, cout = 0;            // real sizes
  int cin_p = 0, cout_p = 0;        // 64-aligned sizes of the split weights
  float out_scale = 1.f;            // constant folded into 2^-k and the bias
  bool up = false;                  // nearest x2 + conv (four parity convs)
  __half* w_hi = nullptr; __half* w_lo = nullptr; float* bias = nullptr; float* wscale = nullptr;
};
}  // namespace cfb

struct cfb_rrdb {
  int in_ch = 3, out_ch = 3, scale = 4, feat = 64, blocks = 23, grow = 32;
  std::mutex mu;
  std::unordered_map<std::string, std::pair<const float*, int64_t>> raw;
  std::vector<cfb::GenConv> convs;      // [blocks*15] dense convs, then conv_body, conv_up1, conv_up2, conv_hr
  float* first_w = nullptr; float* first_b = nullptr;   // conv_first  [tap][cin][64]
  float* last_w = nullptr; float* last_b = nullptr;     // conv_last   [tap][64][4]
  float* slab = nullptr; size_t slab_bytes = 0;
  int device = -1, sm_count = 148;
  bool prepared = false;
};

namespace cfb {

static const float* rrdb_param(cfb_rrdb* n, const std::string& name, int64_t numel) {
  auto it = n->raw.find(name);
  if (it == n->raw.end()) { set_error("missing parameter '" + name + "'"); return nullptr; }
  if (it->second.second != numel) {
    set_error("parameter '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " + std::to_string(numel));
    return nullptr;
  }
  return it->second.first;
}

// zero-padded OIHW copy [cout_p][cin_p][3][3] of a [cout][cin][3][3] weight, then the fp16 hi/lo split of the engine
static int gen_conv_prepare(GenConv& c, const float* w, const float* b, float* pad_scratch, cudaStream_t st) {
  const size_t padn = (size_t)c.cout_p * c.cin_p * 9;
  CFB_CUDA(cudaMemsetAsync(pad_scratch, 0, padn * 4, st));
  CFB_CUDA(cudaMemcpy2DAsync(pad_scratch, (size_t)c.cin_p * 9 * 4, w, (size_t)c.cin * 9 * 4, (size_t)c.cin * 9 * 4, c.cout,
                             cudaMemcpyDeviceToDevice, st));
  if (c.up) CFB_CHECK(tc_split_weights_up4(pad_scratch, c.w_hi, c.w_lo, c.cout_p, c.cin_p, c.wscale, st));
  else CFB_CHECK(tc_split_weights(pad_scratch, c.w_hi, c.w_lo, c.cout_p, c.cin_p, 3, c.wscale, st));
  CFB_CUDA(cudaMemsetAsync(c.bias, 0, (size_t)c.cout_p * 4, st));
  if (b) CFB_CUDA(cudaMemcpyAsync(c.bias, b, (size_t)c.cout * 4, cudaMemcpyDeviceToDevice, st));
  if (c.out_scale != 1.f) {
    CFB_CHECK(scale_scalar(c.wscale + 1, c.out_scale, st));
    CFB_CHECK(scale_vec(c.bias, c.cout, c.out_scale, st));
  }
  return 0;
}

static int rrdb_prepare(cfb_rrdb* n, cudaStream_t st) {
  int dev = 0, major = 0, sms = 148;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CFB_REQUIRE(major == 10, "RRDBNet: the tcgen05 engine needs an sm_100 device (there is no other path)");
  CFB_CHECK(async_status_init(st));
  n->convs.clear();
  const int us = n->scale == 2 ? 2 : (n->scale == 1 ? 4 : 1);
  const int cin_first = n->in_ch * us * us;
  for (int b = 0; b < n->blocks; ++b)
    for (int r = 1; r <= 3; ++r)
      for (int k = 1; k <= 5; ++k) {
        GenConv c;
        c.name = "body." + std::to_string(b) + ".rdb" + std::to_string(r) + ".conv" + std::to_string(k);
        c.cin = n->feat + (k - 1) * n->grow; c.cout = k == 5 ? n->feat : n->grow;
        c.out_scale = k == 5 ? 0.2f : 1.f;
        n->convs.push_back(c);
      }
  for (const char* nm : {"conv_body", "conv_up1", "conv_up2", "conv_hr"}) {
    GenConv c; c.name = nm; c.cin = n->feat; c.cout = n->feat; c.up = (c.name == "conv_up1" || c.name == "conv_up2");
    n->convs.push_back(c);
  }
  size_t total = 0, padmax = 0;
  for (GenConv& c : n->convs) {
    c.cin_p = (c.cin + 63) / 64 * 64; c.cout_p = (c.cout + 63) / 64 * 64;
    const size_t wn = (size_t)c.cout_p * c.cin_p * (c.up ? 16 : 9);
    total += 2 * align256(wn * 2) + align256((size_t)c.cout_p * 4) + 256;
    padmax = std::max(padmax, (size_t)c.cout_p * c.cin_p * 9 * 4);
  }
  total += align256(padmax) + align256((size_t)9 * cin_first * 64 * 4) + 256 + align256((size_t)9 * 64 * 4 * 4) + 256;
  if (n->slab && (n->device != dev || n->slab_bytes < total)) {
    if (n->device != dev && n->device >= 0) { cudaSetDevice(n->device); cudaFree(n->slab); cudaSetDevice(dev); }
    else cudaFree(n->slab);
    n->slab = nullptr; n->slab_bytes = 0;
  }
  if (!n->slab) { CFB_CUDA(cudaMalloc((void**)&n->slab, total)); n->slab_bytes = total; }
  n->device = dev; n->sm_count = sms;
  char* p = (char*)n->slab;
  auto take = [&](size_t bytes) { char* r = p; p += align256(bytes); return r; };
  float* pad_scratch = (float*)take(padmax);
  for (GenConv& c : n->convs) {
    const size_t wn = (size_t)c.cout_p * c.cin_p * (c.up ? 16 : 9);
    c.w_hi = (__half*)take(wn * 2); c.w_lo = (__half*)take(wn * 2);
    c.bias = (float*)take((size_t)c.cout_p * 4); c.wscale = (float*)take(8);
    const float* w = rrdb_param(n, c.name + ".weight", (int64_t)c.cout * c.cin * 9);
    const float* b = rrdb_param(n, c.name + ".bias", c.cout);
    if (!w || !b) return 1;
    CFB_CHECK(gen_conv_prepare(c, w, b, pad_scratch, st));
  }
  {
    const float* w = rrdb_param(n, "conv_first.weight", (int64_t)n->feat * cin_first * 9);
    const float* b = rrdb_param(n, "conv_first.bias", n->feat);
    if (!w || !b) return 1;
    n->first_w = (float*)take((size_t)9 * cin_first * 64 * 4); n->first_b = (float*)take(256);
    CFB_CHECK(relayout_oihw_to_tck(w, n->first_w, n->feat, cin_first, 3, st));
    CFB_CUDA(cudaMemcpyAsync(n->first_b, b, (size_t)n->feat * 4, cudaMemcpyDeviceToDevice, st));
  }
  {
    const float* w = rrdb_param(n, "conv_last.weight", (int64_t)n->out_ch * n->feat * 9);
    const float* b = rrdb_param(n, "conv_last.bias", n->out_ch);
    if (!w || !b) return 1;
    n->last_w = (float*)take((size_t)9 * 64 * 4 * 4); n->last_b = (float*)take(256);
    CFB_CHECK(relayout_thin_out(w, n->last_w, n->out_ch, st));
    CFB_CUDA(cudaMemcpyAsync(n->last_b, b, (size_t)n->out_ch * 4, cudaMemcpyDeviceToDevice, st));
  }
  CFB_CUDA(cudaStreamSynchronize(st));
  n->prepared = true;
  return 0;
}

struct GenLaunch {          // one generalised conv: src window -> destination slice
  const GenConv* c; const float* in; int in_pitch; int H, W; int N;
  float* out; int out_pitch, out_c0; int act;
  const float* res = nullptr; int res_pitch = 0; const float* res2 = nullptr; int res2_pitch = 0; float post = 1.f;
  int pad_mode = 0; bool sub = false;
};
static int gen_conv(const GenLaunch& g, int sm_count, cudaStream_t st) {
  ConvArgs a;
  a.in = g.in; a.N = g.N; a.H = g.H; a.W = g.W; a.Cin = g.c->cin_p;
  a.Ho = g.c->up ? 2 * g.H : g.H; a.Wo = g.c->up ? 2 * g.W : g.W; a.Cout = g.c->cout_p; a.ksize = 3;
  a.mode = g.c->up ? CONV_UP : CONV_SAME;
  a.wgt_hi = g.c->w_hi; a.wgt_lo = g.c->w_lo; a.wscale_inv = g.c->wscale + 1; a.bias = g.c->bias;
  a.residual = g.res; a.out_act = g.act; a.out = g.out;
  a.skip_prep = true; a.xform = true; a.gen = true;
  a.in_pitch = g.in_pitch; a.pad_mode = g.pad_mode; a.subsample = g.sub;
  a.out_pitch = g.out_pitch; a.out_c0 = g.out_c0; a.cout_valid = g.c->cout;
  a.res_pitch = g.res_pitch; a.residual2 = g.res2; a.res2_pitch = g.res2_pitch; a.post_scale = g.post;
  return conv_tc(a, nullptr, sm_count, st);
}

static size_t rrdb_ws_bytes(const cfb_rrdb* n, int N, int H, int W) {
  const int us = n->scale == 2 ? 2 : (n->scale == 1 ? 4 : 1);
  const size_t px = (size_t)N * (H / us) * (W / us);
  // F (64) + three dense buffers (192) + body (64) at low resolution; up1 (64 @2x); up2 + hr (64 @4x)
  return (px * (64 + 3 * 192 + 64) + px * 4 * 64 + px * 16 * 64 * 2) * sizeof(float) + 8 * 1024;
}

static int rrdb_forward(cfb_rrdb* n, const float* x, float* out, int N, int H, int W, void* ws, int64_t ws_bytes, cudaStream_t st) {
  CFB_REQUIRE(n->prepared, "cfb_rrdb_prepare has not been called");
  int dev = -1;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_REQUIRE(dev == n->device, "RRDBNet was prepared on another CUDA device");
  CFB_CHECK(async_status_check("cfb_rrdb_forward"));
  const int us = n->scale == 2 ? 2 : (n->scale == 1 ? 4 : 1);
  CFB_REQUIRE(H % us == 0 && W % us == 0, "RRDBNet: H and W must be multiples of the pixel-unshuffle factor (arch_util.py:202)");
  if (N == 0 || H == 0 || W == 0) return 0;
  CFB_REQUIRE((size_t)ws_bytes >= rrdb_ws_bytes(n, N, H, W), "workspace too small (cfb_rrdb_workspace_bytes)");
  const int h = H / us, w = W / us;
  const size_t px = (size_t)N * h * w;
  float* p = (float*)(((uintptr_t)ws + 1023) / 1024 * 1024);
  float* F = p; p += px * 64;
  float* D[3]; for (int i = 0; i < 3; ++i) { D[i] = p; p += px * 192; }
  float* Bd = p; p += px * 64;
  float* U1 = p; p += px * 4 * 64;
  float* U2 = p; p += px * 16 * 64;
  float* HR = p; p += px * 16 * 64;
  // dense buffers start at zero: every channel a conv window can touch is finite from the first launch on
  CFB_CUDA(cudaMemsetAsync(D[0], 0, px * 192 * 3 * sizeof(float), st));
  CFB_CHECK(conv_thin_in(x, n->first_w, n->first_b, F, N, h, w, n->in_ch, us, 0, 64, 0, st));
  CFB_CHECK(conv_thin_in(x, n->first_w, n->first_b, D[0], N, h, w, n->in_ch, us, 0, 192, 0, st));
  int X = 0, Y = 1;            // x of the current RRDB lives in D[X]; D[2] is the middle buffer
  const int Z = 2;
  for (int b = 0; b < n->blocks; ++b) {
    const int src[3] = {X, Y, Z}, dst[3] = {Y, Z, Y};
    for (int r = 0; r < 3; ++r) {
      float* S = D[src[r]];
      for (int k = 0; k < 5; ++k) {
        const GenConv& c = n->convs[(b * 3 + r) * 5 + k];
        GenLaunch g{&c, S, 192, h, w, N, nullptr, 192, 0, OUT_NONE};
        if (k < 4) { g.out = S; g.out_c0 = 64 + 32 * k; g.act = OUT_LRELU; }           // x_{k+1} = lrelu(conv_k(cat(x, x1..xk)))
        else {
          g.out = D[dst[r]]; g.out_c0 = 0; g.res = S; g.res_pitch = 192;               // x5 * 0.2 + x   (rrdbnet_arch.py:40)
          if (r == 2) { g.res2 = D[X]; g.res2_pitch = 192; g.post = 0.2f; }            // out * 0.2 + x  (rrdbnet_arch.py:63)
        }
        CFB_CHECK(gen_conv(g, n->sm_count, st));
      }
    }
    std::swap(X, Y);
  }
  const size_t nb = (size_t)n->blocks * 15;
  { GenLaunch g{&n->convs[nb + 0], D[X], 192, h, w, N, Bd, 64, 0, OUT_NONE}; g.res = F; g.res_pitch = 64;   // feat + conv_body(body(feat))
    CFB_CHECK(gen_conv(g, n->sm_count, st)); }
  { GenLaunch g{&n->convs[nb + 1], Bd, 64, h, w, N, U1, 64, 0, OUT_LRELU}; CFB_CHECK(gen_conv(g, n->sm_count, st)); }
  { GenLaunch g{&n->convs[nb + 2], U1, 64, 2 * h, 2 * w, N, U2, 64, 0, OUT_LRELU}; CFB_CHECK(gen_conv(g, n->sm_count, st)); }
  { GenLaunch g{&n->convs[nb + 3], U2, 64, 4 * h, 4 * w, N, HR, 64, 0, OUT_LRELU}; CFB_CHECK(gen_conv(g, n->sm_count, st)); }
  CFB_CHECK(conv_thin_out(HR, n->last_w, n->last_b, out, N, 4 * h, 4 * w, n->out_ch, 0, st));
  return 0;
}

}  // namespace cfb

// =========================================================================================================
// ParseNet: the face-parsing network the paste-back step runs on every restored face
//   /root/reference/facelib/parsing/parsenet.py:140-194 (constructor :142-186, forward :188-194), caller
//   /root/reference/facelib/utils/face_restoration_helper.py:457-487.
// Every ConvLayer = ReflectionPad2d(1) + 3x3 conv (+ eval-mode BatchNorm, folded into the weights at prepare) (+ LeakyReLU 0.2).
// The 64..256-channel convs run on the generalised fused-transform tcgen05 engine: reflection padding is produced inside the
// kernel (border pixels of the halo patch are copies of patch pixels), 'down' layers (stride 2) keep the even positions of the
// stride-1 result, 'up' layers (nearest x2 + reflection pad) are four parity convs on the low-resolution tensor with replicate
// padding, and the residual sums `identity + res` / `feat + body(feat)` are epilogue residuals.
// =========================================================================================================
namespace cfb {
struct PnBlock { int kind; int cin, cout; GenConv sc, c1, c2; };     // kind: 0 none, 1 down, 2 up
}
struct cfb_parsenet {
  int in_size = 512, out_size = 512, min_feat = 32, base_ch = 64, parsing_ch = 19, res_depth = 10, ch_min = 32, ch_max = 256;
  std::mutex mu;
  std::unordered_map<std::string, std::pair<const float*, int64_t>> raw;
  std::vector<cfb::PnBlock> blocks;       // encoder[1:], body, decoder in order
  int n_enc = 0, n_body = 0, n_dec = 0, head_ch = 64;
  float *first_w = nullptr, *first_b = nullptr, *mask_w = nullptr, *mask_b = nullptr, *img_w = nullptr, *img_b = nullptr;
  float* slab = nullptr; size_t slab_bytes = 0;
  int device = -1, sm_count = 148;
  bool prepared = false;
  cfb::Arena arena;
};
namespace cfb {

static const float* pn_param(cfb_parsenet* n, const std::string& name, int64_t numel) {
  auto it = n->raw.find(name);
  if (it == n->raw.end()) { set_error("missing parameter '" + name + "'"); return nullptr; }
  if (it->second.second != numel) {
    set_error("parameter '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " + std::to_string(numel));
    return nullptr;
  }
  return it->second.first;
}

static int pn_build(cfb_parsenet* n) {       // ParseNet.__init__  parsenet.py:151-186
  auto clip = [&](int x) { return std::max(n->ch_min, std::min(x, n->ch_max)); };
  const int mfs = std::min(n->in_size, n->min_feat);
  const int down = (int)std::lround(std::floor(std::log2((double)(n->in_size / mfs))));
  const int up = (int)std::lround(std::floor(std::log2((double)(n->out_size / mfs))));
  n->blocks.clear();
  int head = n->base_ch;
  auto add = [&](const std::string& p, int kind, int cin, int cout) {
    PnBlock b; b.kind = kind; b.cin = cin; b.cout = cout;
    b.sc.name = p + ".shortcut_func"; b.c1.name = p + ".conv1"; b.c2.name = p + ".conv2";
    b.sc.cin = cin; b.sc.cout = cout; b.sc.up = kind == 2;
    b.c1.cin = cin; b.c1.cout = cout; b.c1.up = kind == 2;
    b.c2.cin = cout; b.c2.cout = cout;
    n->blocks.push_back(b);
  };
  for (int i = 0; i < down; ++i) { add("encoder." + std::to_string(i + 1), 1, clip(head), clip(head * 2)); head *= 2; }
  n->n_enc = down;
  for (int i = 0; i < n->res_depth; ++i) add("body." + std::to_string(i), 0, clip(head), clip(head));
  n->n_body = n->res_depth;
  for (int i = 0; i < up; ++i) { add("decoder." + std::to_string(i), 2, clip(head), clip(head / 2)); head /= 2; }
  n->n_dec = up;
  n->head_ch = clip(head);
  CFB_REQUIRE(n->base_ch == 64 && n->head_ch == 64, "ParseNet: built for base_ch = 64 and a 64-channel head (in_size == out_size)");
  for (const PnBlock& b : n->blocks) {
    CFB_REQUIRE(b.cin % 64 == 0 && b.cout % 64 == 0, "ParseNet: channel counts must be multiples of 64");
    CFB_REQUIRE(b.kind != 0 || b.cin == b.cout, "ParseNet: a body block with a channel change (conv shortcut) is not built");
  }
  CFB_REQUIRE(n->parsing_ch >= 1 && n->parsing_ch <= 20, "ParseNet: at most 20 parsing classes");
  return 0;
}

static int pn_prepare(cfb_parsenet* n, cudaStream_t st) {
  int dev = 0, major = 0, sms = 148;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CFB_REQUIRE(major == 10, "ParseNet: the tcgen05 engine needs an sm_100 device (there is no other path)");
  CFB_CHECK(async_status_init(st));
  CFB_CHECK(pn_build(n));
  size_t total = 0, padmax = 0;
  auto acct = [&](GenConv& c) {
    c.cin_p = c.cin; c.cout_p = c.cout;
    const size_t wn = (size_t)c.cout_p * c.cin_p * (c.up ? 16 : 9);
    total += 2 * align256(wn * 2) + align256((size_t)c.cout_p * 4) + 256;
    padmax = std::max(padmax, (size_t)c.cout_p * c.cin_p * 9 * 4);
  };
  for (PnBlock& b : n->blocks) { if (b.kind) acct(b.sc); acct(b.c1); acct(b.c2); }
  total += 2 * align256(padmax) + align256(1024) + align256((size_t)27 * 64 * 4) + 256 + 2 * (align256((size_t)9 * 64 * 20 * 4) + 256);
  if (n->slab && (n->device != dev || n->slab_bytes < total)) {
    if (n->device != dev && n->device >= 0) { cudaSetDevice(n->device); cudaFree(n->slab); cudaSetDevice(dev); }
    else cudaFree(n->slab);
    n->slab = nullptr; n->slab_bytes = 0;
  }
  if (!n->slab) { CFB_CUDA(cudaMalloc((void**)&n->slab, total)); n->slab_bytes = total; }
  n->device = dev; n->sm_count = sms;
  char* p = (char*)n->slab;
  auto take = [&](size_t bytes) { char* r = p; p += align256(bytes); return r; };
  float* pad_scratch = (float*)take(padmax);
  float* fold_w = (float*)take(padmax);
  float* fold_b = (float*)take(1024);
  auto prep = [&](GenConv& c, bool bn) -> int {
    const size_t wn = (size_t)c.cout_p * c.cin_p * (c.up ? 16 : 9);
    c.w_hi = (__half*)take(wn * 2); c.w_lo = (__half*)take(wn * 2);
    c.bias = (float*)take((size_t)c.cout_p * 4); c.wscale = (float*)take(8);
    const float* w = pn_param(n, c.name + ".conv2d.weight", (int64_t)c.cout * c.cin * 9);
    if (!w) return 1;
    if (!bn) {
      const float* b = pn_param(n, c.name + ".conv2d.bias", c.cout);
      if (!b) return 1;
      return gen_conv_prepare(c, w, b, pad_scratch, st);
    }
    const float* g = pn_param(n, c.name + ".norm.norm.weight", c.cout);
    const float* be = pn_param(n, c.name + ".norm.norm.bias", c.cout);
    const float* mu = pn_param(n, c.name + ".norm.norm.running_mean", c.cout);
    const float* var = pn_param(n, c.name + ".norm.norm.running_var", c.cout);
    if (!g || !be || !mu || !var) return 1;
    CFB_REQUIRE(c.cout <= 256, "ParseNet: more than 256 channels");
    CFB_CHECK(fold_bn(w, g, be, mu, var, 1e-5f, fold_w, fold_b, c.cout, c.cin * 9, st));      // nn.BatchNorm2d default eps
    return gen_conv_prepare(c, fold_w, fold_b, pad_scratch, st);
  };
  for (PnBlock& b : n->blocks) {
    if (b.kind) CFB_CHECK(prep(b.sc, false));
    CFB_CHECK(prep(b.c1, true));
    CFB_CHECK(prep(b.c2, true));
  }
  {
    const float* w = pn_param(n, "encoder.0.conv2d.weight", (int64_t)64 * 3 * 9);
    const float* b = pn_param(n, "encoder.0.conv2d.bias", 64);
    if (!w || !b) return 1;
    n->first_w = (float*)take((size_t)27 * 64 * 4); n->first_b = (float*)take(256);
    CFB_CHECK(relayout_oihw_to_tck(w, n->first_w, 64, 3, 3, st));
    CFB_CUDA(cudaMemcpyAsync(n->first_b, b, 64 * 4, cudaMemcpyDeviceToDevice, st));
  }
  for (int which = 0; which < 2; ++which) {
    const std::string nm = which ? "out_img_conv" : "out_mask_conv";
    const int co = which ? 3 : n->parsing_ch;
    const float* w = pn_param(n, nm + ".conv2d.weight", (int64_t)co * 64 * 9);
    const float* b = pn_param(n, nm + ".conv2d.bias", co);
    if (!w || !b) return 1;
    float* wd = (float*)take((size_t)9 * 64 * 20 * 4);
    float* bd = (float*)take(256);
    CFB_CHECK(relayout_thin_out(w, wd, co, st));
    CFB_CUDA(cudaMemcpyAsync(bd, b, (size_t)co * 4, cudaMemcpyDeviceToDevice, st));
    if (which) { n->img_w = wd; n->img_b = bd; } else { n->mask_w = wd; n->mask_b = bd; }
  }
  CFB_CUDA(cudaStreamSynchronize(st));
  n->prepared = true;
  return 0;
}

static int pn_forward(cfb_parsenet* n, const float* x, float* out_mask, float* out_img, int N, int H, int W, void* ws, int64_t ws_bytes,
                      cudaStream_t st, bool dry) {
  CFB_REQUIRE(dry || n->prepared, "cfb_parsenet_prepare has not been called");
  if (!dry) {
    int dev = -1;
    CFB_CUDA(cudaGetDevice(&dev));
    CFB_REQUIRE(dev == n->device, "ParseNet was prepared on another CUDA device");
    CFB_CHECK(async_status_check("cfb_parsenet_forward"));
  }
  const int div = 1 << n->n_enc;
  CFB_REQUIRE(H % div == 0 && W % div == 0 && H >= 2 * div && W >= 2 * div, "ParseNet: H and W must be multiples of 2^down_steps");
  if (N == 0) return 0;
  Arena& ar = n->arena;
  ar.reset(ws, (size_t)ws_bytes, dry);
  auto alloc = [&](float** p, size_t elems) -> int {
    *p = (float*)ar.alloc(elems * sizeof(float));
    CFB_REQUIRE(*p != nullptr, "workspace too small (cfb_parsenet_workspace_bytes)");
    return 0;
  };
  float* t = nullptr;
  int h = H, w = W, c = 64;
  CFB_CHECK(alloc(&t, (size_t)N * h * w * 64));
  if (!dry) CFB_CHECK(conv_thin_in(x, n->first_w, n->first_b, t, N, h, w, 3, 1, 1, 64, 0, st));
  float* feat = nullptr;           // encoder output, added back after the body (parsenet.py:190)
  for (size_t i = 0; i < n->blocks.size(); ++i) {
    const PnBlock& b = n->blocks[i];
    CFB_REQUIRE(b.cin == c, "ParseNet: channel plan mismatch");
    const bool last_body = (int)i == n->n_enc + n->n_body - 1 && n->n_body > 0;
    if ((int)i == n->n_enc) feat = t;
    float *s = nullptr, *c1 = nullptr, *o = nullptr;
    int ho = h, wo = w;
    if (b.kind == 1) { ho = h / 2; wo = w / 2; } else if (b.kind == 2) { ho = 2 * h; wo = 2 * w; }
    const int h1 = b.kind == 2 ? ho : h, w1 = b.kind == 2 ? wo : w;          // resolution of conv1's output
    if (b.kind) {
      CFB_CHECK(alloc(&s, (size_t)N * ho * wo * b.cout));
      GenLaunch g{&b.sc, t, b.cin, h, w, N, s, b.cout, 0, OUT_NONE};
      g.pad_mode = b.kind == 2 ? 2 : 1; g.sub = b.kind == 1;
      if (!dry) CFB_CHECK(gen_conv(g, n->sm_count, st));
    }
    CFB_CHECK(alloc(&c1, (size_t)N * h1 * w1 * b.cout));
    {
      GenLaunch g{&b.c1, t, b.cin, h, w, N, c1, b.cout, 0, OUT_LRELU};
      g.pad_mode = b.kind == 2 ? 2 : 1;
      if (!dry) CFB_CHECK(gen_conv(g, n->sm_count, st));
    }
    CFB_CHECK(alloc(&o, (size_t)N * ho * wo * b.cout));
    {
      GenLaunch g{&b.c2, c1, b.cout, h1, w1, N, o, b.cout, 0, OUT_NONE};
      g.pad_mode = 1; g.sub = b.kind == 1;
      g.res = b.kind ? s : t; g.res_pitch = b.cout;
      if (last_body) { g.res2 = feat; g.res2_pitch = b.cout; g.post = 1.f; }      // x = feat + body(feat)
      if (!dry) CFB_CHECK(gen_conv(g, n->sm_count, st));
    }
    ar.release(c1);
    if (s) ar.release(s);
    if (t != feat) ar.release(t);
    if (last_body && feat) { ar.release(feat); feat = nullptr; }
    t = o; h = ho; w = wo; c = b.cout;
  }
  if (n->n_body == 0) feat = nullptr;
  CFB_REQUIRE(c == 64, "ParseNet: head must have 64 channels");
  if (!dry) {
    CFB_CHECK(conv_thin_out(t, n->mask_w, n->mask_b, out_mask, N, h, w, n->parsing_ch, 1, st));
    if (out_img) CFB_CHECK(conv_thin_out(t, n->img_w, n->img_b, out_img, N, h, w, 3, 1, st));
  }
  ar.release(t);
  return 0;
}

}  // namespace cfb

// =========================================================================================================
// C ABI
// =========================================================================================================
#define API_BEGIN try {
#define API_END(ret)                                                             \
  } catch (const std::exception& e) { cfb::set_error(std::string("exception: ") + e.what()); return ret; } \
  catch (...) { cfb::set_error("unknown exception"); return ret; }

extern "C" {

int cfb_version(void) { return CFB_VERSION; }
const char* cfb_last_error(void) { return cfb::last_error().c_str(); }

int cfb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  API_BEGIN
  int dev = 0;
  CFB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  CFB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
  API_END(1)
}

cfb_net* cfb_net_create(const cfb_config* cfg) {
  API_BEGIN
  if (!cfg) { cfb::set_error("cfb_net_create: NULL config"); return nullptr; }
  cfb_net* n = new cfb_net();
  n->cfg = *cfg;
  n->convs.reserve(512);
  if (cfb::build_plan(n) != 0) { delete n; return nullptr; }
  int dev = 0, major = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  } else {
    cudaGetLastError();
  }
  n->sm_count = sms > 0 ? sms : 148;
  n->tc_ok = (major == 10);
  n->device = dev;
  return n;
  API_END(nullptr)
}

void cfb_net_destroy(cfb_net* n) {
  if (!n) return;
  if (n->slab) {
    int cur = -1;
    const bool sw = cudaGetDevice(&cur) == cudaSuccess && n->device >= 0 && cur != n->device;
    if (sw) cudaSetDevice(n->device);
    cudaFree(n->slab);
    if (sw) cudaSetDevice(cur);
  }
  delete n;
}

cfb_rrdb* cfb_rrdb_create(int32_t num_in_ch, int32_t num_out_ch, int32_t scale, int32_t num_feat, int32_t num_block, int32_t num_grow_ch) {
  API_BEGIN
  if (num_feat != 64 || num_grow_ch != 32 || num_in_ch < 1 || num_in_ch > 3 || num_out_ch < 1 || num_out_ch > 4 || num_block < 1 ||
      !(scale == 1 || scale == 2 || scale == 4)) {
    cfb::set_error("cfb_rrdb_create: built for num_feat=64, num_grow_ch=32, <=3 image channels, scale 1/2/4");
    return nullptr;
  }
  cfb_rrdb* n = new cfb_rrdb();
  n->in_ch = num_in_ch; n->out_ch = num_out_ch; n->scale = scale; n->feat = num_feat; n->blocks = num_block; n->grow = num_grow_ch;
  return n;
  API_END(nullptr)
}
void cfb_rrdb_destroy(cfb_rrdb* n) {
  if (!n) return;
  if (n->slab) {
    int cur = -1;
    const bool sw = cudaGetDevice(&cur) == cudaSuccess && n->device >= 0 && cur != n->device;
    if (sw) cudaSetDevice(n->device);
    cudaFree(n->slab);
    if (sw) cudaSetDevice(cur);
  }
  delete n;
}
int cfb_rrdb_set_param(cfb_rrdb* n, const char* name, const float* dev_ptr, int64_t numel) {
  API_BEGIN
  CFB_REQUIRE(n && name && dev_ptr, "cfb_rrdb_set_param: NULL argument");
  std::lock_guard<std::mutex> lk(n->mu);
  n->raw[name] = {dev_ptr, numel};
  n->prepared = false;
  return 0;
  API_END(1)
}
int cfb_rrdb_prepare(cfb_rrdb* n, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n, "cfb_rrdb_prepare: NULL net");
  std::lock_guard<std::mutex> lk(n->mu);
  return cfb::rrdb_prepare(n, (cudaStream_t)stream);
  API_END(1)
}
int64_t cfb_rrdb_workspace_bytes(cfb_rrdb* n, int32_t batch, int32_t h, int32_t w) {
  if (!n || batch < 0 || h < 0 || w < 0) return -1;
  return (int64_t)cfb::rrdb_ws_bytes(n, batch, h, w);
}
int cfb_rrdb_forward(cfb_rrdb* n, const float* x, float* out, int32_t batch, int32_t h, int32_t w, void* workspace,
                     int64_t workspace_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n && (batch == 0 || (x && out && workspace)), "cfb_rrdb_forward: NULL argument");
  std::lock_guard<std::mutex> lk(n->mu);
  return cfb::rrdb_forward(n, x, out, batch, h, w, workspace, workspace_bytes, (cudaStream_t)stream);
  API_END(1)
}

cfb_parsenet* cfb_parsenet_create(int32_t in_size, int32_t out_size, int32_t min_feat_size, int32_t base_ch, int32_t parsing_ch,
                                  int32_t res_depth, int32_t ch_min, int32_t ch_max) {
  API_BEGIN
  cfb_parsenet* n = new cfb_parsenet();
  n->in_size = in_size; n->out_size = out_size; n->min_feat = min_feat_size; n->base_ch = base_ch; n->parsing_ch = parsing_ch;
  n->res_depth = res_depth; n->ch_min = ch_min; n->ch_max = ch_max;
  if (in_size < 1 || out_size < 1 || min_feat_size < 1 || cfb::pn_build(n) != 0) { delete n; return nullptr; }
  return n;
  API_END(nullptr)
}
void cfb_parsenet_destroy(cfb_parsenet* n) {
  if (!n) return;
  if (n->slab) {
    int cur = -1;
    const bool sw = cudaGetDevice(&cur) == cudaSuccess && n->device >= 0 && cur != n->device;
    if (sw) cudaSetDevice(n->device);
    cudaFree(n->slab);
    if (sw) cudaSetDevice(cur);
  }
  delete n;
}
int cfb_parsenet_set_param(cfb_parsenet* n, const char* name, const float* dev_ptr, int64_t numel) {
  API_BEGIN
  CFB_REQUIRE(n && name && dev_ptr, "cfb_parsenet_set_param: NULL argument");
  std::lock_guard<std::mutex> lk(n->mu);
  n->raw[name] = {dev_ptr, numel};
  n->prepared = false;
  return 0;
  API_END(1)
}
int cfb_parsenet_prepare(cfb_parsenet* n, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n, "cfb_parsenet_prepare: NULL net");
  std::lock_guard<std::mutex> lk(n->mu);
  return cfb::pn_prepare(n, (cudaStream_t)stream);
  API_END(1)
}
int64_t cfb_parsenet_workspace_bytes(cfb_parsenet* n, int32_t batch, int32_t h, int32_t w) {
  API_BEGIN
  if (!n) { cfb::set_error("cfb_parsenet_workspace_bytes: NULL net"); return -1; }
  std::lock_guard<std::mutex> lk(n->mu);
  if (cfb::pn_forward(n, (const float*)0x1000, (float*)0x1000, (float*)0x1000, batch, h, w, nullptr, 0, nullptr, true) != 0) return -1;
  return (int64_t)n->arena.high() + 4096;
  API_END(-1)
}
int cfb_parsenet_forward(cfb_parsenet* n, const float* x, float* out_mask, float* out_img, int32_t batch, int32_t h, int32_t w,
                         void* workspace, int64_t workspace_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n && (batch == 0 || (x && out_mask && workspace)), "cfb_parsenet_forward: NULL argument");
  std::lock_guard<std::mutex> lk(n->mu);
  return cfb::pn_forward(n, x, out_mask, out_img, batch, h, w, workspace, workspace_bytes, (cudaStream_t)stream, false);
  API_END(1)
}
int cfb_parse_argmax(const float* logits_nchw, uint8_t* classes, uint8_t* mask, int32_t batch, int32_t channels, int64_t hw, void* stream) {
  API_BEGIN
  CFB_REQUIRE(logits_nchw && (classes || mask), "cfb_parse_argmax: NULL argument");
  return cfb::parse_argmax(logits_nchw, classes, mask, batch, channels, hw, (cudaStream_t)stream);
  API_END(1)
}

int64_t cfb_conv2d_gen_workspace_bytes(int32_t cin, int32_t cout) {
  const size_t cin_p = (size_t)(cin + 63) / 64 * 64, cout_p = (size_t)(cout + 63) / 64 * 64;
  return (int64_t)(align256(cout_p * cin_p * 9 * 4) + 2 * align256(cout_p * cin_p * 16 * 2) + align256(cout_p * 4) + 256 + 4096);
}

int cfb_conv2d_gen_nhwc(const float* in, int32_t in_pitch, const float* weight_oihw, const float* bias, float* out,
                        int32_t out_pitch, int32_t out_c0, int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout,
                        int32_t upsample, int32_t pad_mode, int32_t subsample, int32_t out_act, const float* residual,
                        int32_t res_pitch, const float* residual2, int32_t res2_pitch, float post_scale, void* workspace,
                        int64_t workspace_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(in && weight_oihw && out && workspace, "cfb_conv2d_gen_nhwc: NULL argument");
  CFB_REQUIRE(workspace_bytes >= cfb_conv2d_gen_workspace_bytes(cin, cout), "cfb_conv2d_gen_nhwc: workspace too small");
  CFB_REQUIRE(cin >= 1 && cout >= 1 && cout % 4 == 0, "cfb_conv2d_gen_nhwc: cout must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  CFB_CHECK(cfb::async_status_init(st));
  int dev = 0, sms = 148;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cfb::GenConv c;
  c.cin = cin; c.cout = cout; c.cin_p = (cin + 63) / 64 * 64; c.cout_p = (cout + 63) / 64 * 64; c.up = upsample != 0;
  char* p = (char*)(((uintptr_t)workspace + 255) / 256 * 256);
  float* pad = (float*)p; p += align256((size_t)c.cout_p * c.cin_p * 9 * 4);
  const size_t wn = (size_t)c.cout_p * c.cin_p * (c.up ? 16 : 9);
  c.w_hi = (__half*)p; p += align256(wn * 2);
  c.w_lo = (__half*)p; p += align256(wn * 2);
  c.bias = (float*)p; p += align256((size_t)c.cout_p * 4);
  c.wscale = (float*)p;
  CFB_CHECK(cfb::gen_conv_prepare(c, weight_oihw, bias, pad, st));
  cfb::GenLaunch g{&c, in, in_pitch, h, w, n, out, out_pitch, out_c0, out_act};
  g.res = residual; g.res_pitch = res_pitch; g.res2 = residual2; g.res2_pitch = res2_pitch; g.post = post_scale;
  g.pad_mode = pad_mode; g.sub = subsample != 0;
  return cfb::gen_conv(g, sms, st);
  API_END(1)
}

int cfb_check_async_status(void) {
  API_BEGIN
  return cfb::async_status_check("cfb_check_async_status");
  API_END(1)
}

int cfb_debug_set_wait_limit(int64_t cycles) {
  API_BEGIN
  CFB_REQUIRE(cycles > 0, "cfb_debug_set_wait_limit: cycles must be positive");
  CFB_CHECK(cfb::async_status_init(nullptr));
  {
    std::lock_guard<std::mutex> lk(cfb::g_status_mu);
    cfb::g_wait_limit_cycles = cycles;
  }
  int dev = 0;
  CFB_CUDA(cudaGetDevice(&dev));
  unsigned* dptr = nullptr;
  CFB_CUDA(cudaHostGetDevicePointer((void**)&dptr, cfb::g_status_words, 0));
  return cfb::tc_bind_status_word(dptr + dev, cycles);
  API_END(1)
}

int cfb_debug_inject_fault(int32_t kind) {
  API_BEGIN
  return cfb::tc_inject_fault(kind);
  API_END(1)
}
int cfb_debug_set_stamps(int64_t* stamps) {
  API_BEGIN
  return cfb::tc_set_stamps((long long*)stamps);
  API_END(1)
}

int cfb_net_set_param(cfb_net* n, const char* name, const float* dev_ptr, int64_t numel) {
  API_BEGIN
  CFB_REQUIRE(n && name && dev_ptr, "cfb_net_set_param: NULL argument");
  std::lock_guard<std::mutex> lk(n->mu);
  n->raw[name] = {dev_ptr, numel};
  n->prepared = false;
  return 0;
  API_END(1)
}

int cfb_net_prepare(cfb_net* n, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n, "cfb_net_prepare: NULL net");
  std::lock_guard<std::mutex> lk(n->mu);
  return cfb::prepare(n, (cudaStream_t)stream);
  API_END(1)
}

int64_t cfb_workspace_bytes(cfb_net* n, int32_t batch) {
  API_BEGIN
  if (!n) { cfb::set_error("cfb_workspace_bytes: NULL net"); return -1; }
  std::lock_guard<std::mutex> lk(n->mu);
  {
    auto it = n->ws_memo.find(batch);      // 16 host-side dry-run plans per miss: remember the answer per batch size
    if (it != n->ws_memo.end()) return it->second;
  }
  // The arena is first-fit, so the peak depends on the exact allocation sequence, which the call flags change (fusion on
  // or off, AdaIN, code_only, caller-provided logits or not): size for the worst of all of them (host-only dry runs).
  size_t high = 0;
  if (n->cfg.kind == 1) {
    for (int m = 0; m < 16; ++m) {
      const float w = (m & 1) ? 1.f : 0.f;
      float* lg = (m & 8) ? (float*)0x1000 : nullptr;
      if (cfb::codeformer_forward_impl(n, (const float*)0x1000, (float*)0x1000, lg, (float*)0x1000, nullptr, batch, w, (m >> 1) & 1,
                                       (m >> 2) & 1, nullptr, 0, nullptr, true) != 0)
        return -1;
      if (n->arena.high() > high) high = n->arena.high();
    }
  } else {
    for (int m = 0; m < 2; ++m) {
      if (cfb::vqae_forward_impl(n, (const float*)0x1000, (float*)0x1000, m ? (int64_t*)0x1000 : nullptr, m ? (float*)0x1000 : nullptr,
                                 nullptr, batch, nullptr, 0, nullptr, true) != 0)
        return -1;
      if (n->arena.high() > high) high = n->arena.high();
    }
  }
  n->ws_memo[batch] = (int64_t)high + 4096;
  return (int64_t)high + 4096;
  API_END(-1)
}

int64_t cfb_last_launch_count(cfb_net* n) { return n ? n->last_launches : 0; }

int cfb_net_set_engine(cfb_net* n, int32_t engine) {
  API_BEGIN
  CFB_REQUIRE(n && engine >= 0 && engine <= 2, "cfb_net_set_engine: engine must be 0 (auto), 1 (fp32) or 2 (tcgen05)");
  std::lock_guard<std::mutex> lk(n->mu);
  n->engine = engine;
  n->ws_memo.clear();
  return 0;
  API_END(1)
}

int cfb_net_capture(cfb_net* n, const char* stage, float* dst, int64_t capacity) {
  API_BEGIN
  CFB_REQUIRE(n && stage, "cfb_net_capture: NULL argument");
  std::lock_guard<std::mutex> lk(n->mu);
  if (dst) n->captures[stage] = {dst, capacity};
  else n->captures.erase(stage);
  n->ws_memo.clear();
  return 0;
  API_END(1)
}

int cfb_codeformer_forward(cfb_net* n, const float* x, float* out, float* logits, float* lq_feat, int64_t* top_idx,
                           int32_t batch, float w, int32_t adain, int32_t code_only, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n, "cfb_codeformer_forward: NULL net");
  if (batch == 0) return 0;
  CFB_REQUIRE(x, "cfb_codeformer_forward: NULL input");
  std::lock_guard<std::mutex> lk(n->mu);   // two caller threads may share one net (web-demos/hugging_face/app.py:282)
  const int64_t before = cfb::launch_count();
  const int rc = cfb::codeformer_forward_impl(n, x, out, logits, lq_feat, top_idx, batch, w, adain, code_only, workspace,
                                              workspace_bytes, (cudaStream_t)stream, false);
  n->last_launches = cfb::launch_count() - before;
  return rc;
  API_END(1)
}

int cfb_codeformer_forward_u8(cfb_net* n, const uint8_t* faces_bgr, uint8_t* restored_bgr, float* logits, float* lq_feat,
                              int64_t* top_idx, int32_t batch, float w, int32_t adain, void* workspace,
                              int64_t workspace_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n, "cfb_codeformer_forward_u8: NULL net");
  if (batch == 0) return 0;
  CFB_REQUIRE(faces_bgr && restored_bgr, "cfb_codeformer_forward_u8: NULL image pointer");
  std::lock_guard<std::mutex> lk(n->mu);
  const int64_t before = cfb::launch_count();
  const int rc = cfb::codeformer_forward_impl(n, nullptr, nullptr, logits, lq_feat, top_idx, batch, w, adain, 0, workspace,
                                              workspace_bytes, (cudaStream_t)stream, false, faces_bgr, restored_bgr);
  n->last_launches = cfb::launch_count() - before;
  return rc;
  API_END(1)
}

int cfb_codeformer_restore_host(cfb_net* n, const uint8_t* faces_host, uint8_t* restored_host, int32_t batch, float w,
                                int32_t adain, void* dev_scratch, int64_t dev_scratch_bytes, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n && faces_host && restored_host && dev_scratch, "cfb_codeformer_restore_host: NULL argument");
  CFB_REQUIRE(dev_scratch_bytes >= cfb_host_io_bytes(n, batch), "dev_scratch too small (cfb_host_io_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t img = (size_t)batch * 3 * n->cfg.img_size * n->cfg.img_size;
  uint8_t* din = (uint8_t*)dev_scratch;
  uint8_t* dout = din + (img + 1023) / 1024 * 1024;
  CFB_CUDA(cudaMemcpyAsync(din, faces_host, img, cudaMemcpyHostToDevice, st));
  CFB_CHECK(cfb_codeformer_forward_u8(n, din, dout, nullptr, nullptr, nullptr, batch, w, adain, workspace, workspace_bytes, stream));
  CFB_CUDA(cudaMemcpyAsync(restored_host, dout, img, cudaMemcpyDeviceToHost, st));
  CFB_CUDA(cudaStreamSynchronize(st));
  return cfb::async_status_check("cfb_codeformer_restore_host");
  API_END(1)
}

int cfb_u8_to_input(const uint8_t* img_bgr_hwc, float* x_nchw, int32_t n, int32_t hw, void* stream) {
  API_BEGIN
  CFB_REQUIRE((img_bgr_hwc && x_nchw) || n == 0, "cfb_u8_to_input: NULL argument");
  return cfb::u8_to_input(img_bgr_hwc, x_nchw, n, hw, (cudaStream_t)stream);
  API_END(1)
}
int cfb_output_to_u8(const float* x_nchw, uint8_t* img_bgr_hwc, int32_t n, int32_t hw, void* stream) {
  API_BEGIN
  CFB_REQUIRE((img_bgr_hwc && x_nchw) || n == 0, "cfb_output_to_u8: NULL argument");
  return cfb::output_to_u8(x_nchw, img_bgr_hwc, n, hw, (cudaStream_t)stream);
  API_END(1)
}

int64_t cfb_host_io_bytes(cfb_net* n, int32_t batch) {
  if (!n) return -1;
  const cfb_config& c = n->cfg;
  const int64_t img = (int64_t)batch * 3 * c.img_size * c.img_size * 4;
  const int64_t lat = (int64_t)batch * c.latent_size;
  return 2 * (img + 1024) + lat * c.codebook_size * 4 + lat * c.emb_dim * 4 + 4096;
}

int cfb_codeformer_forward_host(cfb_net* n, const float* x_host, float* out_host, float* logits_host, float* lq_host,
                                int32_t batch, float w, int32_t adain, void* dev_scratch, int64_t dev_scratch_bytes,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n && x_host && out_host && dev_scratch, "cfb_codeformer_forward_host: NULL argument");
  CFB_REQUIRE(dev_scratch_bytes >= cfb_host_io_bytes(n, batch), "dev_scratch too small (cfb_host_io_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  const cfb_config& c = n->cfg;
  const size_t img = (size_t)batch * 3 * c.img_size * c.img_size * 4;
  const size_t lat = (size_t)batch * c.latent_size;
  char* p = (char*)dev_scratch;
  float* dx = (float*)p; p += (img + 1023) / 1024 * 1024;
  float* dout = (float*)p; p += (img + 1023) / 1024 * 1024;
  float* dlog = (float*)p; p += lat * c.codebook_size * 4;
  float* dlq = (float*)p;
  CFB_CUDA(cudaMemcpyAsync(dx, x_host, img, cudaMemcpyHostToDevice, st));
  CFB_CHECK(cfb_codeformer_forward(n, dx, dout, dlog, dlq, nullptr, batch, w, adain, 0, workspace, workspace_bytes, stream));
  CFB_CUDA(cudaMemcpyAsync(out_host, dout, img, cudaMemcpyDeviceToHost, st));
  if (logits_host) CFB_CUDA(cudaMemcpyAsync(logits_host, dlog, lat * c.codebook_size * 4, cudaMemcpyDeviceToHost, st));
  if (lq_host) CFB_CUDA(cudaMemcpyAsync(lq_host, dlq, lat * c.emb_dim * 4, cudaMemcpyDeviceToHost, st));
  CFB_CUDA(cudaStreamSynchronize(st));
  return cfb::async_status_check("cfb_codeformer_forward_host");
  API_END(1)
}

int cfb_vqae_forward(cfb_net* n, const float* x, float* out, int64_t* idx, float* stats, float* min_encodings,
                     int32_t batch, void* workspace, int64_t workspace_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(n, "cfb_vqae_forward: NULL net");
  if (batch == 0) return 0;
  CFB_REQUIRE(x && out, "cfb_vqae_forward: NULL argument");
  std::lock_guard<std::mutex> lk(n->mu);
  const int64_t before = cfb::launch_count();
  const int rc = cfb::vqae_forward_impl(n, x, out, idx, stats, min_encodings, batch, workspace, workspace_bytes,
                                        (cudaStream_t)stream, false);
  n->last_launches = cfb::launch_count() - before;
  return rc;
  API_END(1)
}

static bool vq_tc_args(cfb::ConvArgs& a, int batch, int h, int w, int dim, int codes) {
  a = cfb::ConvArgs();
  a.N = batch; a.H = h; a.W = w; a.Cin = dim; a.Ho = h; a.Wo = w; a.Cout = codes; a.ksize = 1; a.mode = cfb::CONV_SAME;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return major == 10 && cfb::tc_supported(a);
}

int64_t cfb_vq_workspace_bytes(int32_t batch, int32_t hw, int32_t dim, int32_t codes) {
  const int64_t T = (int64_t)batch * hw;
  const int64_t tok = 2 * ((T * dim * 4 + 1023) / 1024 * 1024);
  const int64_t simt = (int64_t)cfb::vq_workspace_bytes((int)T, dim, codes);
  // tensor-core path (16x16-style latents): split codebook + operand planes + the [T,K] dot products
  const int64_t wsplit = 2 * (int64_t)align256((size_t)codes * dim * 2) + 256;
  const int64_t planes = 2 * ((T * dim * 2 + 1023) / 1024 * 1024);
  const int64_t dots = (T * codes * 4 + 1023) / 1024 * 1024;
  const int64_t tc = wsplit + planes + dots + (int64_t)cfb::vq_select_workspace_bytes((int)T, codes) + 4096;
  return tok + (simt > tc ? simt : tc) + 8192;
}

int cfb_vq_nearest(const float* z, const float* codebook, int32_t batch, int32_t h, int32_t w, int32_t dim, int32_t codes,
                   float beta, float* z_q, int64_t* idx, float* stats, float* min_encodings, void* workspace,
                   int64_t workspace_bytes, void* stream) {
  API_BEGIN
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t T = (int64_t)batch * h * w;
  if (T == 0) return 0;                     // empty batch: nothing to do (stats are left untouched)
  CFB_REQUIRE(z && codebook && z_q && idx && stats && workspace, "cfb_vq_nearest: NULL argument");
  CFB_REQUIRE(workspace_bytes >= cfb_vq_workspace_bytes(batch, h * w, dim, codes), "cfb_vq_nearest: workspace too small");
  const size_t tb = ((size_t)T * dim * 4 + 1023) / 1024 * 1024;
  char* p = (char*)workspace;
  float* zt = (float*)p; p += tb;
  float* zqt = (float*)p; p += tb;
  CFB_CHECK(cfb::nchw_to_nhwc(z, zt, batch, dim, h * w, st));
  cfb::ConvArgs a;
  if (vq_tc_args(a, batch, h, w, dim, codes)) {
    __half* whi = (__half*)p; p += align256((size_t)codes * dim * 2);
    __half* wlo = (__half*)p; p += align256((size_t)codes * dim * 2);
    float* wsc = (float*)p; p += 256;
    p = (char*)(((uintptr_t)p + 1023) / 1024 * 1024);
    void* planes = p; p += 2 * (((size_t)T * dim * 2 + 1023) / 1024 * 1024);
    float* dots = (float*)p; p += ((size_t)T * codes * 4 + 1023) / 1024 * 1024;
    a.in = zt; a.out = dots; a.wgt_hi = whi; a.wgt_lo = wlo; a.wscale_inv = wsc + 1;
    int dev = 0, sms = 148;
    CFB_CUDA(cudaGetDevice(&dev));
    CFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CFB_CHECK(cfb::tc_split_weights(codebook, whi, wlo, codes, dim, 1, wsc, st));
    CFB_CHECK(cfb::conv_tc(a, planes, sms, st));
    CFB_CHECK(cfb::vq_select_from_dots(zt, codebook, dots, (int)T, dim, codes, beta, idx, zqt, stats, min_encodings, p, st));
  } else {
    CFB_CHECK(cfb::vq_nearest(zt, codebook, (int)T, dim, codes, beta, idx, zqt, stats, min_encodings, p, st));
  }
  CFB_CHECK(cfb::nhwc_to_nchw(zqt, z_q, batch, dim, h * w, st));
  return 0;
  API_END(1)
}

// ---- VectorQuantizer.forward, fused path (BASELINE config 3): 4 launches on NCHW tensors ------------------------------------
int32_t cfb_vq_fast_supported(int32_t batch, int32_t h, int32_t w, int32_t dim, int32_t codes) {
  cfb::ConvArgs a;
  if (!vq_tc_args(a, batch, h, w, dim, codes)) return 0;
  return (codes % 128 == 0 && dim % 64 == 0 && dim <= 352 && (h * w) % 128 == 0 && batch >= 0) ? 1 : 0;
}
int64_t cfb_vq_prepared_bytes(int32_t codes, int32_t dim) {
  // split codebook (hi, lo), weight scale, |e|^2, and the self-cleaning code histogram + ticket of the one-kernel path
  return (int64_t)(2 * align256((size_t)codes * dim * 2) + 256 + 2 * align256((size_t)codes * 4) + 256);
}
int cfb_vq_prepare(const float* codebook, int32_t codes, int32_t dim, void* prepared, int64_t prepared_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(codebook && prepared && prepared_bytes >= cfb_vq_prepared_bytes(codes, dim), "cfb_vq_prepare: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  char* p = (char*)prepared;
  __half* whi = (__half*)p; p += align256((size_t)codes * dim * 2);
  __half* wlo = (__half*)p; p += align256((size_t)codes * dim * 2);
  float* wsc = (float*)p; p += 256;
  float* e2 = (float*)p;
  CFB_CHECK(cfb::tc_split_weights(codebook, whi, wlo, codes, dim, 1, wsc, st));
  CFB_CHECK(cfb::vq_e2(codebook, e2, codes, dim, st));
  CFB_CUDA(cudaMemsetAsync((char*)e2 + align256((size_t)codes * 4), 0, align256((size_t)codes * 4) + 256, st));   // histogram + ticket
  return 0;
  API_END(1)
}
int64_t cfb_vq_fast_workspace_bytes(int32_t batch, int32_t hw, int32_t dim, int32_t codes) {
  const int64_t T = (int64_t)batch * hw;
  const int64_t planes = 2 * ((T * dim * 2 + 1023) / 1024 * 1024);
  const int64_t ncand = 2 * (codes / 128);
  return planes + align256(T * 4) + align256(T * ncand * 8) + align256((T / 128 + 1) * (codes / 128) * 8 * 8) + align256((T / 32 + 1) * 8) +
         align256((size_t)codes * 4) + 8192;
}
int cfb_vq_nearest_fast(const float* z, const float* codebook, const void* prepared, int32_t batch, int32_t h, int32_t w, int32_t dim,
                        int32_t codes, float beta, float* z_q, int64_t* idx, float* stats, float* min_encodings, void* workspace,
                        int64_t workspace_bytes, void* stream) {
  API_BEGIN
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = h * w;
  const int64_t T = (int64_t)batch * HW;
  if (T == 0) return 0;
  CFB_REQUIRE(z && codebook && prepared && z_q && idx && stats && workspace, "cfb_vq_nearest_fast: NULL argument");
  CFB_REQUIRE(cfb_vq_fast_supported(batch, h, w, dim, codes), "cfb_vq_nearest_fast: shape not on the fused path (use cfb_vq_nearest)");
  CFB_REQUIRE(workspace_bytes >= cfb_vq_fast_workspace_bytes(batch, HW, dim, codes), "cfb_vq_nearest_fast: workspace too small");
  CFB_CHECK(cfb::async_status_init(st));
  const char* q = (const char*)prepared;
  const __half* whi = (const __half*)q; q += align256((size_t)codes * dim * 2);
  const __half* wlo = (const __half*)q; q += align256((size_t)codes * dim * 2);
  const float* wsc = (const float*)q; q += 256;
  const float* e2 = (const float*)q;
  char* p = (char*)(((uintptr_t)workspace + 1023) / 1024 * 1024);
  if (cfb::vq_fused_supported(batch, dim, HW, codes) && !(getenv("CFB_VQ_FUSED") && atoi(getenv("CFB_VQ_FUSED")) == 0)) {
    // ONE kernel (conv_tc.cu: vq_fused_kernel): histogram / ticket live in the prepared buffer (zero between calls)
    unsigned* fh = (unsigned*)((char*)e2 + align256((size_t)codes * 4));
    unsigned* ticket = (unsigned*)((char*)fh + align256((size_t)codes * 4));
    // CFB_VQ_TIMING=1 (tools/gpu_r2_vq.sh): clock64 stamps of the phase boundaries per CTA, 4 KB into the workspace
    long long* dbg = (getenv("CFB_VQ_TIMING") && atoi(getenv("CFB_VQ_TIMING"))) ? (long long*)(p + 4096) : nullptr;
    CFB_CHECK(cfb::vq_fused(z, codebook, whi, wlo, wsc + 1, e2, fh, ticket, (double*)p, batch, dim, HW, codes, beta, z_q, idx, stats, st, dbg));
    if (min_encodings) CFB_CHECK(cfb::onehot_from_idx(idx, min_encodings, (int)T, codes, st));
    return 0;
  }
  void* planes = p; p += 2 * (((size_t)T * dim * 2 + 1023) / 1024 * 1024);
  float* z2 = (float*)p; p += align256((size_t)T * 4);
  const int ncand = 2 * (codes / 128);
  float2* cand = (float2*)p; p += align256((size_t)T * ncand * 8);
  const int n_d = (int)(T / 128) * (codes / 128) * 8;
  double* dpart = (double*)p; p += align256((size_t)(T / 128 + 1) * (codes / 128) * 8 * 8);
  const int n_se = (int)(T / 32);
  double* separt = (double*)p; p += align256((size_t)(T / 32 + 1) * 8);
  unsigned* hist = (unsigned*)p;
  int dev = 0, sms = 148;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CFB_CHECK(cfb::vq_prep_nchw(z, planes, z2, hist, batch, dim, HW, codes, st));
  cfb::ConvArgs a;
  a.N = batch; a.H = h; a.W = w; a.Cin = dim; a.Ho = h; a.Wo = w; a.Cout = codes; a.ksize = 1; a.mode = cfb::CONV_SAME;
  a.wgt_hi = whi; a.wgt_lo = wlo; a.wscale_inv = wsc + 1; a.skip_prep = true;
  a.vq_e2 = e2; a.vq_z2 = z2; a.vq_cand = cand; a.vq_dpart = dpart;
  CFB_CHECK(cfb::conv_tc(a, planes, sms, st));
  CFB_CHECK(cfb::vq_select_cand(z, codebook, cand, ncand, batch, dim, HW, codes, idx, z_q, separt, hist, st));
  CFB_CHECK(cfb::vq_final2(separt, n_se, dpart, n_d, hist, (int)T, dim, codes, beta, stats, st));
  if (min_encodings) CFB_CHECK(cfb::onehot_from_idx(idx, min_encodings, (int)T, codes, st));
  return 0;
  API_END(1)
}

int cfb_codebook_lookup(const int64_t* idx, const float* codebook, int32_t batch, int32_t h, int32_t w, int32_t dim,
                        int32_t codes, float* z_q, void* stream) {
  API_BEGIN
  CFB_REQUIRE(idx && codebook && z_q, "cfb_codebook_lookup: NULL argument");
  // gather token-major then transpose in place is not possible; gather straight into NCHW order instead
  // (small: B*256 tokens) via a temporary-free two-step is avoided by a strided gather kernel:
  cudaStream_t st = (cudaStream_t)stream;
  const int T = batch * h * w;
  if (T == 0) return 0;
  float* tmp = nullptr;
  CFB_CUDA(cudaMallocAsync((void**)&tmp, (size_t)T * dim * 4, st));
  int rc = cfb::gather_rows(idx, codebook, tmp, T, codes, dim, st);
  if (rc == 0) rc = cfb::nhwc_to_nchw(tmp, z_q, batch, dim, h * w, st);
  cudaFreeAsync(tmp, st);
  return rc;
  API_END(1)
}

int64_t cfb_conv2d_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t ksize, int32_t mode) {
  cfb::ConvArgs a;
  a.N = n; a.H = h; a.W = w; a.Cin = cin; a.Cout = cout; a.ksize = ksize; a.mode = mode;
  a.Ho = mode == cfb::CONV_DOWN ? h / 2 : (mode == cfb::CONV_UP ? h * 2 : h);
  a.Wo = mode == cfb::CONV_DOWN ? w / 2 : (mode == cfb::CONV_UP ? w * 2 : w);
  const size_t wn = (size_t)cout * cin * ksize * ksize;
  const size_t wsplit = mode == cfb::CONV_UP ? (size_t)16 * cout * cin : wn;
  return (int64_t)(align256(wn * 4) + 2 * align256(wsplit * 2) + 256 + cfb::tc_scratch_bytes(a) + 8192);
}

int cfb_conv2d_nhwc(const float* in, const float* weight_oihw, const float* bias, float* out, int32_t n, int32_t h,
                    int32_t w, int32_t cin, int32_t cout, int32_t ksize, int32_t mode, const float* in_scale,
                    const float* in_shift, int32_t in_act, const float* residual, int32_t out_act, int32_t engine,
                    void* workspace, int64_t workspace_bytes, void* stream) {
  API_BEGIN
  CFB_REQUIRE(in && weight_oihw && out && workspace, "cfb_conv2d_nhwc: NULL argument");
  CFB_CHECK(cfb::async_status_init(nullptr));
  CFB_REQUIRE(workspace_bytes >= cfb_conv2d_workspace_bytes(n, h, w, cin, cout, ksize, mode), "cfb_conv2d_nhwc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  cfb::ConvArgs a;
  a.in = in; a.N = n; a.H = h; a.W = w; a.Cin = cin; a.Cout = cout; a.ksize = ksize; a.mode = mode;
  a.Ho = mode == cfb::CONV_DOWN ? h / 2 : (mode == cfb::CONV_UP ? h * 2 : h);
  a.Wo = mode == cfb::CONV_DOWN ? w / 2 : (mode == cfb::CONV_UP ? w * 2 : w);
  a.bias = bias; a.in_scale = in_scale; a.in_shift = in_shift; a.in_act = in_act; a.residual = residual;
  a.out_act = out_act; a.out = out;
  const size_t wn = (size_t)cout * cin * ksize * ksize;
  char* p = (char*)workspace;
  float* wf = (float*)p; p += align256(wn * 4);
  const size_t wsplit = mode == cfb::CONV_UP ? (size_t)16 * cout * cin : wn;
  __half* whi = (__half*)p; p += align256(wsplit * 2);
  __half* wlo = (__half*)p; p += align256(wsplit * 2);
  float* wsc = (float*)p; p += 256;
  p = (char*)(((uintptr_t)p + 1023) / 1024 * 1024);
  a.wgt_f32 = wf; a.wgt_hi = whi; a.wgt_lo = wlo; a.wscale_inv = wsc + 1;
  bool use_tc = engine == 2;
  if (engine == 0) {
    int dev = 0, major = 0;
    CFB_CUDA(cudaGetDevice(&dev));
    CFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    use_tc = major == 10 && cfb::tc_supported(a);
  }
  if (use_tc) {
    CFB_REQUIRE(cfb::tc_supported(a), "cfb_conv2d_nhwc: shape not supported by the tcgen05 engine");
    int dev = 0, sms = 148;
    CFB_CUDA(cudaGetDevice(&dev));
    CFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (mode == cfb::CONV_UP) CFB_CHECK(cfb::tc_split_weights_up4(weight_oihw, whi, wlo, cout, cin, wsc, st));
    else CFB_CHECK(cfb::tc_split_weights(weight_oihw, whi, wlo, cout, cin, ksize, wsc, st));
    // same choice as the network runtime (Fwd::conv): a GroupNorm-affine (+SiLU) input goes through the in-kernel operand
    // transform wherever the engine has it, so the kernel tests exercise the path the forward ships
    if (in_scale && in_shift && cfb::tc_can_xform(a)) { a.xform = true; a.skip_prep = true; }
    CFB_CHECK(cfb::conv_tc(a, p, sms, st));
  } else {
    CFB_CHECK(cfb::relayout_oihw_to_tck(weight_oihw, wf, cout, cin, ksize, st));
    CFB_CHECK(cfb::conv_f32(a, st));
  }
  return 0;
  API_END(1)
}

int cfb_debug_umma_probe(const void* a_f16, int32_t rows_a, const void* b_f16, const int32_t* cfg_dev, int32_t ncfg, float* out,
                         void* stream) {
  API_BEGIN
  return cfb::umma_probe(a_f16, rows_a, b_f16, cfg_dev, ncfg, out, (cudaStream_t)stream);
  API_END(1)
}

int cfb_debug_umma_rate(int32_t n, int32_t nacc, int32_t reps, int64_t* out_dev, int32_t ctas, void* stream) {
  API_BEGIN
  return cfb::umma_rate(n, nacc, reps, (long long*)out_dev, ctas, (cudaStream_t)stream);
  API_END(1)
}

int cfb_debug_umma_pair(int32_t n, int32_t reps, float* vals_dev, int64_t* info_dev, int32_t ctas, void* stream) {
  API_BEGIN
  return cfb::umma_pair(n, reps, vals_dev, (long long*)info_dev, ctas, (cudaStream_t)stream);
  API_END(1)
}

int cfb_debug_time_conv(const float* in, const float* weight_oihw, float* out, int32_t n, int32_t h, int32_t w, int32_t cin,
                        int32_t cout, int32_t ksize, int32_t mode, int32_t reps, void* workspace, int64_t workspace_bytes,
                        void* stream, const float* in_scale, const float* in_shift, int32_t in_act, float* ms_per_launch) {
  API_BEGIN
  CFB_REQUIRE(in && weight_oihw && out && workspace && ms_per_launch && reps > 0, "cfb_debug_time_conv: bad argument");
  CFB_REQUIRE(workspace_bytes >= cfb_conv2d_workspace_bytes(n, h, w, cin, cout, ksize, mode), "cfb_debug_time_conv: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  cfb::ConvArgs a;
  a.in = in; a.N = n; a.H = h; a.W = w; a.Cin = cin; a.Cout = cout; a.ksize = ksize; a.mode = mode;
  a.Ho = mode == cfb::CONV_DOWN ? h / 2 : (mode == cfb::CONV_UP ? h * 2 : h);
  a.Wo = mode == cfb::CONV_DOWN ? w / 2 : (mode == cfb::CONV_UP ? w * 2 : w);
  a.out = out;
  CFB_REQUIRE(cfb::tc_supported(a), "cfb_debug_time_conv: shape not on the tcgen05 engine");
  const size_t wn = (size_t)cout * cin * ksize * ksize;
  const size_t wsplit = mode == cfb::CONV_UP ? (size_t)16 * cout * cin : wn;
  char* p = (char*)workspace + align256(wn * 4);
  __half* whi = (__half*)p; p += align256(wsplit * 2);
  __half* wlo = (__half*)p; p += align256(wsplit * 2);
  float* wsc = (float*)p; p += 256;
  p = (char*)(((uintptr_t)p + 1023) / 1024 * 1024);
  a.wgt_hi = whi; a.wgt_lo = wlo; a.wscale_inv = wsc + 1;
  int dev = 0, sms = 148;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (mode == cfb::CONV_UP) CFB_CHECK(cfb::tc_split_weights_up4(weight_oihw, whi, wlo, cout, cin, wsc, st));
  else CFB_CHECK(cfb::tc_split_weights(weight_oihw, whi, wlo, cout, cin, ksize, wsc, st));
  CFB_CHECK(cfb::async_status_init(nullptr));
  a.in_scale = in_scale; a.in_shift = in_shift; a.in_act = in_act;
  // a GroupNorm-affine (+SiLU) input: the kernel timed is the one the forward ships -- fused operand transform where the
  // engine has it (all-in: no separate prep pass exists), otherwise prep once outside the timed region
  if (in_scale && in_shift && cfb::tc_can_xform(a)) { a.xform = true; a.skip_prep = true; }
  CFB_CHECK(cfb::conv_tc(a, p, sms, st));          // operand prep + one warm-up launch
  a.skip_prep = true;
  for (int i = 0; i < 2; ++i) CFB_CHECK(cfb::conv_tc(a, p, sms, st));
  cudaEvent_t e0, e1;
  CFB_CUDA(cudaEventCreate(&e0));
  CFB_CUDA(cudaEventCreate(&e1));
  CFB_CUDA(cudaEventRecord(e0, st));               // events on the launching stream: only the conv kernel is between them
  for (int i = 0; i < reps; ++i) CFB_CHECK(cfb::conv_tc(a, p, sms, st));
  CFB_CUDA(cudaEventRecord(e1, st));
  CFB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  CFB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_per_launch = ms / reps;
  return 0;
  API_END(1)
}

int64_t cfb_gn_workspace_bytes(int32_t n, int32_t hw, int32_t c) { return (int64_t)cfb::gn_workspace_bytes(n, hw, c) + 256; }
int cfb_group_norm_coef(const float* x, const float* gamma, const float* beta, float* scale, float* shift, int32_t n,
                        int32_t hw, int32_t c, int32_t groups, float eps, void* workspace, int64_t workspace_bytes,
                        void* stream) {
  API_BEGIN
  CFB_REQUIRE(x && gamma && beta && scale && shift && workspace, "cfb_group_norm_coef: NULL argument");
  CFB_REQUIRE(workspace_bytes >= (int64_t)cfb::gn_workspace_bytes(n, hw, c), "cfb_group_norm_coef: workspace too small");
  return cfb::gn_coef(x, gamma, beta, scale, shift, n, hw, c, groups, eps, workspace, (cudaStream_t)stream);
  API_END(1)
}
int cfb_affine_act(const float* x, const float* scale, const float* shift, float* y, int32_t n, int32_t hw, int32_t c,
                   int32_t act, void* stream) {
  API_BEGIN
  return cfb::affine_act(x, scale, shift, y, n, hw, c, act, (cudaStream_t)stream);
  API_END(1)
}
int cfb_attention(const float* q, const float* k, const float* v, float* out, int32_t batch, int32_t tokens, int32_t heads,
                  int32_t d, int32_t q_pitch, int32_t k_pitch, int32_t v_pitch, int32_t o_pitch, float scale, void* stream) {
  API_BEGIN
  return cfb::attention(q, k, v, out, batch, tokens, heads, d, q_pitch, k_pitch, v_pitch, o_pitch, scale, (cudaStream_t)stream);
  API_END(1)
}
int cfb_layer_norm(const float* x, const float* gamma, const float* beta, float* y, float* y2, const float* pos,
                   int32_t pos_rows, int32_t rows, int32_t c, void* stream) {
  API_BEGIN
  return cfb::layer_norm(x, gamma, beta, y, y2, pos, pos_rows, rows, c, (cudaStream_t)stream);
  API_END(1)
}
int cfb_adain_nhwc(const float* content, const float* style, float* out, int32_t batch, int32_t hw, int32_t c, void* stream) {
  API_BEGIN
  return cfb::adain_nhwc(content, style, out, batch, hw, c, (cudaStream_t)stream);
  API_END(1)
}
int cfb_nchw_to_nhwc(const float* in, float* out, int32_t n, int32_t c, int32_t hw, void* stream) {
  API_BEGIN
  return cfb::nchw_to_nhwc(in, out, n, c, hw, (cudaStream_t)stream);
  API_END(1)
}
int cfb_nhwc_to_nchw(const float* in, float* out, int32_t n, int32_t c, int32_t hw, void* stream) {
  API_BEGIN
  return cfb::nhwc_to_nchw(in, out, n, c, hw, (cudaStream_t)stream);
  API_END(1)
}

}  // extern "C"
