// C-ABI of ladiff_b200 (include/ladiff_b200.h): handle, strict loader, load-time folds, stage plans.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "codec_ops.cuh"
#include "common.cuh"
#include "fold.cuh"
#include "unet_ops.cuh"

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
void ladiff_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* ladiff_last_error(void) { return g_err; }
bool ladiff_pdl_enabled() {
  static const bool on = getenv("LADIFF_NO_PDL") == nullptr;
  return on;
}
extern "C" int32_t ladiff_abi_version(void) { return 2; }
extern "C" const char* ladiff_act_dtype(void) { return LADIFF_DTYPE_NAME; }

#define TRY(expr)              \
  do {                         \
    int _rc = (expr);          \
    if (_rc != 0) return _rc;  \
  } while (0)

namespace {

const int kTimesteps = 1000;
const int kBins = 1024;
const int kMults[5] = {1, 2, 2, 4, 4};

struct KeySpec {
  std::string name;
  std::vector<int64_t> shape;
  bool alias;   // diffusion.model.* duplicates diff_model.* (presence/shape checked, data not copied)
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

// ---- folded codec parameters
struct ConvW { const float* w = nullptr; const float* wt = nullptr; const float* bias = nullptr; int Cout = 0, Cin = 0, K = 0, stride = 1; CodecTcWeights tcw; int precise = 0; };
struct ConvTrW { const float* w2 = nullptr; const float* wt = nullptr; const float* bias = nullptr; int Cin = 0, Cout = 0, s = 1; CodecTcWeights tcw; };
struct ResBlockW { ConvW c1, c2, sc; };
struct LstmW { int H = 0, layers = 0; const float* wih[4]; const float* wih_t[4]; const float* whh[4]; const float* bias[4]; CodecTcWeights tcw[4]; int precise = 0; };
struct EncoderW { ConvW first, last; std::vector<ResBlockW> rb; std::vector<ConvW> down; LstmW lstm; };
struct DecoderW { ConvW first, last; std::vector<ConvTrW> up; std::vector<ResBlockW> rb; LstmW lstm; };

// ---- folded UNet parameters
enum ConvKind { CK_PLAIN = TC_KIND_PLAIN, CK_DOWN = TC_KIND_DOWN, CK_UP = TC_KIND_UP };
struct PackedConv {
  h16* w = nullptr; const float* bias = nullptr;
  int CoutV = 0, Cin = 0, K = 1, Ktot = 0, kind = CK_PLAIN;
  int split_m = 0;     // > 0: rows [split_m, CoutV) are a fused 1x1 res_conv of the same input (second output)
  CUtensorMap tmW;
};
struct ResnetW {
  PackedConv c1, c2; bool has_res = false;
  const float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
  long long film_off = 0; int Cin = 0, Cout = 0;
};
struct AttnW { PackedConv qkv, out; const float* norm_g = nullptr; const float* out_g = nullptr; int C = 0; };
struct UNetW {
  int dims[6];
  PackedConv init, finalc, down[5], up[5];
  ResnetW d[5][2], u[5][2], mid1, mid2, fin;
  AttnW da[5], ua[5], mida;
  float* film = nullptr; long long film_stride = 0;
  DdpmTables tb;
  // host copies of the schedule buffers the samplers read (ddpm_loss.py:140-168), device pointers for the per-clip kernels
  std::vector<float> h_recip, h_recipm1, h_c1, h_c2, h_logvar, h_ac;
  const float *d_sqrt_ac = nullptr, *d_sqrt_1m_ac = nullptr, *d_p2w = nullptr;
  std::vector<ConvTrW> cond_up;
};

struct Bump {
  uintptr_t base; size_t off = 0;
  explicit Bump(void* p) : base((uintptr_t)p) {}
  template <class T> T* get(size_t n) {
    off = align_up(base + off, 1024) - base;    // absolute 1 KB alignment whatever the caller's base alignment
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

struct UnetBufs {
  h16 *xin, *FC, *CA[5], *CB[5], *X[6], *tY, *tH, *tO, *tR, *tA, *qkv, *ao;
  float* eps; float* ctx; float* la_part; int* la_cnt; float2* stats; int* t_dev; float* inv_scale; float* condup; float* condtmp;
  int stats_slots;
};

typedef std::function<int(cudaStream_t)> Op;
struct Plan {
  void* ws = nullptr; int B = 0, L = 0;
  UnetBufs bufs;
  std::vector<Op> ops;
  std::vector<double> op_flops;          // algorithmic FLOPs of conv ops (2*Cout*Cin*k*Lout*B), 0 for the others
  std::vector<std::string> op_label;
  std::vector<cudaEvent_t> ev;           // profiling: [2*i], [2*i+1] around conv op i; last two around the whole evaluation
  long long launches_per_run = 0;
  long long runs = 0;
  cudaGraphExec_t gexec = nullptr;       // captured evaluation (see run_plan)
  int graph_impl = 0;
  cudaStream_t cap_stream = nullptr;
  ~Plan() {
    for (auto e : ev) cudaEventDestroy(e);
    if (gexec) cudaGraphExecDestroy(gexec);
    if (cap_stream) cudaStreamDestroy(cap_stream);
  }
};

}  // namespace

struct LadiffHandle {
  LadiffConfig cfg;
  std::vector<KeySpec> keys;
  std::map<std::string, int> key_index;
  std::vector<float*> dev;
  std::vector<char> loaded;
  std::vector<void*> owned;
  bool finalized = false;
  int conv_impl = 0;
  int skip_ops = 0;
  int profiling = 0;
  Plan* last_plan = nullptr;
  long long launches = 0;
  unsigned long long clip_offset = 0;   // global index of this handle's clip 0 (in-kernel noise is keyed by the global clip)
  int enc_hop = 1;
  EncoderW enc; DecoderW dec;
  float* embed = nullptr; float* embed_sq = nullptr; float* embed_t = nullptr;
  UNetW un;
  std::vector<Plan*> plans;
  // tile shapes the autotuner picked, by conv signature: every plan of this handle (other workspaces, the second slot of a
  // SynthesisPipeline) reuses them, so all plans of one process sum the GroupNorm partials in the same order -> bitwise equal results
  std::map<std::string, std::vector<int>> tuned;
};

namespace {

typedef LadiffHandle H;

// ------------------------------------------------------------------------------------------------ key layout
// Mirrors ladiffcodec_b200/layout.py (state_dict order of DiffAudioRep, reference srcs/model.py:34-106).
void add_key(H* h, const std::string& n, std::vector<int64_t> s, bool alias = false) {
  h->key_index[n] = (int)h->keys.size();
  h->keys.push_back(KeySpec{n, std::move(s), alias});
}
void keys_wn_conv(H* h, const std::string& p, int cout, int cin, int k) {
  add_key(h, p + ".conv.conv.bias", {cout});
  add_key(h, p + ".conv.conv.weight_g", {cout, 1, 1});
  add_key(h, p + ".conv.conv.weight_v", {cout, cin, k});
}
void keys_wn_convtr(H* h, const std::string& p, int cin, int cout, int k) {
  add_key(h, p + ".convtr.convtr.bias", {cout});
  add_key(h, p + ".convtr.convtr.weight_g", {cin, 1, 1});
  add_key(h, p + ".convtr.convtr.weight_v", {cin, cout, k});
}
void keys_resblock(H* h, const std::string& p, int dim) {
  keys_wn_conv(h, p + ".block.1", dim / 2, dim, 3);
  keys_wn_conv(h, p + ".block.3", dim, dim / 2, 1);
  keys_wn_conv(h, p + ".shortcut", dim, dim, 1);
}
void keys_lstm(H* h, const std::string& p, int dim, int layers) {
  for (int l = 0; l < layers; ++l) {
    const std::string s = std::to_string(l);
    add_key(h, p + ".lstm.weight_ih_l" + s, {4 * dim, dim});
    add_key(h, p + ".lstm.weight_hh_l" + s, {4 * dim, dim});
    add_key(h, p + ".lstm.bias_ih_l" + s, {4 * dim});
    add_key(h, p + ".lstm.bias_hh_l" + s, {4 * dim});
  }
}
std::string M(const std::string& prefix, int i) { return prefix + ".model." + std::to_string(i); }

void keys_unet_resnet(H* h, const std::string& p, int cin, int cout, int td, bool alias) {
  auto A = [&](const std::string& n, std::vector<int64_t> s) { add_key(h, n, std::move(s), alias); };
  A(p + ".mlp.1.weight", {2 * cout, td});
  A(p + ".mlp.1.bias", {2 * cout});
  for (int b = 1; b <= 2; ++b) {
    const std::string q = p + ".block" + std::to_string(b);
    A(q + ".proj.weight", {cout, b == 1 ? cin : cout, 3});
    A(q + ".proj.bias", {cout});
    A(q + ".norm.weight", {cout});
    A(q + ".norm.bias", {cout});
  }
  if (cin != cout) {
    A(p + ".res_conv.weight", {cout, cin, 1});
    A(p + ".res_conv.bias", {cout});
  }
}
void keys_unet_linattn(H* h, const std::string& p, int dim, bool alias) {
  auto A = [&](const std::string& n, std::vector<int64_t> s) { add_key(h, n, std::move(s), alias); };
  A(p + ".fn.fn.to_qkv.weight", {384, dim, 1});
  A(p + ".fn.fn.to_out.0.weight", {dim, 128, 1});
  A(p + ".fn.fn.to_out.0.bias", {dim});
  A(p + ".fn.fn.to_out.1.g", {1, dim, 1});
  A(p + ".fn.norm.g", {1, dim, 1});
}
void keys_unet(H* h, const std::string& pre, bool alias) {
  const LadiffConfig& c = h->cfg;
  auto A = [&](const std::string& n, std::vector<int64_t> s) { add_key(h, n, std::move(s), alias); };
  const int dim = c.diff_dims, inp = c.rep_dims, td = dim * 4, in_ch = inp + c.cond_channels;
  int dims[6]; dims[0] = dim;
  for (int i = 0; i < 5; ++i) dims[i + 1] = dim * kMults[i];
  A(pre + ".init_conv.weight", {dim, in_ch, 7});
  A(pre + ".init_conv.bias", {dim});
  A(pre + ".time_mlp.1.weight", {td, dim});
  A(pre + ".time_mlp.1.bias", {td});
  A(pre + ".time_mlp.3.weight", {td, td});
  A(pre + ".time_mlp.3.bias", {td});
  for (int i = 0; i < 5; ++i) {
    const std::string p = pre + ".downs." + std::to_string(i);
    keys_unet_resnet(h, p + ".0", dims[i], dims[i], td, alias);
    keys_unet_resnet(h, p + ".1", dims[i], dims[i], td, alias);
    keys_unet_linattn(h, p + ".2", dims[i], alias);
    A(p + ".3.weight", {dims[i + 1], dims[i], i < 4 ? 4 : 3});
    A(p + ".3.bias", {dims[i + 1]});
  }
  for (int j = 0; j < 5; ++j) {
    const int i = 4 - j, di = dims[i], dout = dims[i + 1];
    const std::string p = pre + ".ups." + std::to_string(j);
    keys_unet_resnet(h, p + ".0", dout + di, dout, td, alias);
    keys_unet_resnet(h, p + ".1", dout + di, dout, td, alias);
    keys_unet_linattn(h, p + ".2", dout, alias);
    const std::string q = p + (j < 4 ? ".3.1" : ".3");
    A(q + ".weight", {di, dout, 3});
    A(q + ".bias", {di});
  }
  const int mid = dims[5];
  keys_unet_resnet(h, pre + ".mid_block1", mid, mid, td, alias);
  A(pre + ".mid_attn.fn.fn.to_qkv.weight", {384, mid, 1});
  A(pre + ".mid_attn.fn.fn.to_out.weight", {mid, 128, 1});
  A(pre + ".mid_attn.fn.fn.to_out.bias", {mid});
  A(pre + ".mid_attn.fn.norm.g", {1, mid, 1});
  keys_unet_resnet(h, pre + ".mid_block2", mid, mid, td, alias);
  keys_unet_resnet(h, pre + ".final_res_block", dim * 2, dim, td, alias);
  A(pre + ".final_conv.weight", {inp, dim, 1});
  A(pre + ".final_conv.bias", {inp});
  for (int j = 0; j < c.n_upsampling_ratios; ++j) {
    const std::string p = pre + ".upsampling_layers." + std::to_string(j) + ".convtr.convtr";
    A(p + ".weight", {c.cond_channels, c.cond_channels, 2 * c.upsampling_ratios[j]});
    A(p + ".bias", {c.cond_channels});
  }
}

const char* kSchedNames[13] = {"betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                               "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                               "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                               "posterior_mean_coef1", "posterior_mean_coef2", "p2_loss_weight"};

void build_keys(H* h) {
  const LadiffConfig& c = h->cfg;
  const int nf = c.n_filters, dim = c.rep_dims, nr = c.n_enc_ratios;
  {  // encoder, seanet.py:108-151 (ratios reversed)
    int i = 0, mult = 1;
    keys_wn_conv(h, M("encoder", i++), nf, 1, 7);
    for (int r = nr - 1; r >= 0; --r) {
      keys_resblock(h, M("encoder", i++), mult * nf);
      i++;
      keys_wn_conv(h, M("encoder", i++), mult * nf * 2, mult * nf, 2 * c.enc_ratios[r]);
      mult *= 2;
    }
    if (c.lstm_layers) keys_lstm(h, M("encoder", i++), mult * nf, c.lstm_layers);
    i++;
    keys_wn_conv(h, M("encoder", i), dim, mult * nf, 7);
  }
  {  // decoder, seanet.py:201-236
    int i = 0, mult = 1 << nr;
    keys_wn_conv(h, M("decoder", i++), mult * nf, dim, 7);
    if (c.lstm_layers) keys_lstm(h, M("decoder", i++), mult * nf, c.lstm_layers);
    for (int r = 0; r < nr; ++r) {
      i++;
      keys_wn_convtr(h, M("decoder", i++), mult * nf, mult * nf / 2, 2 * c.enc_ratios[r]);
      keys_resblock(h, M("decoder", i++), mult * nf / 2);
      mult /= 2;
    }
    i++;
    keys_wn_conv(h, M("decoder", i), 1, nf, 7);
  }
  if (c.quantization) {
    for (int q = 0; q < c.n_q; ++q) {
      const std::string p = "quantizer.vq.layers." + std::to_string(q) + "._codebook";
      add_key(h, p + ".inited", {1});
      add_key(h, p + ".cluster_size", {kBins});
      add_key(h, p + ".embed", {kBins, dim});
      add_key(h, p + ".embed_avg", {kBins, dim});
    }
  }
  if (c.run_diff) {
    keys_unet(h, "diff_model", false);
    for (int i = 0; i < 13; ++i) add_key(h, std::string("diffusion.") + kSchedNames[i], {kTimesteps});
    keys_unet(h, "diffusion.model", true);
  }
}

// ------------------------------------------------------------------------------------------------ folds
float* Wp(H* h, const std::string& n) {
  auto it = h->key_index.find(n);
  return it == h->key_index.end() ? nullptr : h->dev[it->second];
}
template <class T> int dalloc(H* h, T** p, size_t n) {
  void* q = nullptr;
  LADIFF_CUDA_OK(cudaMalloc(&q, n * sizeof(T)));
  h->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

// hi / lo TF32 planes + tensor map of a K-major weight copy for the tcgen05 codec conv (codec_tc.cu); convs it does not cover keep valid = 0
int prep_codec_tc(H* h, const float* wt, int CinV, int KT, int CoutV, CodecTcWeights* out) {
  if (CinV < 32 || CinV % 4 != 0 || KT > 8 || CoutV < 32) return 0;
  float* planes = nullptr;
  TRY(dalloc(h, &planes, (size_t)2 * CinV * KT * CoutV));
  return codec_tc_prepare(wt, CinV, KT, CoutV, planes, out, 0);
}

int fold_wn_conv(H* h, const std::string& p, int cout, int cin, int k, int stride, ConvW* out) {
  float* w = nullptr;
  TRY(dalloc(h, &w, (size_t)cout * cin * k));
  TRY(weight_norm_fold_launch(Wp(h, p + ".conv.conv.weight_g"), Wp(h, p + ".conv.conv.weight_v"), w, cout, cin * k, 0));
  out->w = w; out->bias = Wp(h, p + ".conv.conv.bias");
  out->Cout = cout; out->Cin = cin; out->K = k; out->stride = stride;
  if (k % stride == 0) {          // K-major copy for the register-tiled kernel
    float* wt = nullptr;
    TRY(dalloc(h, &wt, (size_t)cout * cin * k));
    TRY(conv_w_transpose_launch(w, wt, cout, cin, k, stride, 0, 0, 0));
    out->wt = wt;
    TRY(prep_codec_tc(h, wt, cin * stride, k / stride, cout, &out->tcw));
  }
  return 0;
}
int fold_wn_convtr(H* h, const std::string& p, int cin, int cout, int s, ConvTrW* out) {
  float *w = nullptr, *w2 = nullptr;
  TRY(dalloc(h, &w, (size_t)cin * cout * 2 * s));
  TRY(dalloc(h, &w2, (size_t)cin * cout * 2 * s));
  TRY(weight_norm_fold_launch(Wp(h, p + ".convtr.convtr.weight_g"), Wp(h, p + ".convtr.convtr.weight_v"), w, cin, cout * 2 * s, 0));
  TRY(convtr_pack_launch(w, w2, cin, cout, s, 0));
  out->w2 = w2; out->bias = Wp(h, p + ".convtr.convtr.bias"); out->Cin = cin; out->Cout = cout; out->s = s;
  {
    float* wt = nullptr;
    TRY(dalloc(h, &wt, (size_t)cin * cout * 2 * s));
    TRY(conv_w_transpose_launch(w2, wt, s * cout, cin, 2, 1, s, cout, 0));
    out->wt = wt;
    TRY(prep_codec_tc(h, wt, cin, 2, s * cout, &out->tcw));
  }
  return 0;
}
int fold_resblock(H* h, const std::string& p, int dim, ResBlockW* rb) {
  TRY(fold_wn_conv(h, p + ".block.1", dim / 2, dim, 3, 1, &rb->c1));
  TRY(fold_wn_conv(h, p + ".block.3", dim, dim / 2, 1, 1, &rb->c2));
  TRY(fold_wn_conv(h, p + ".shortcut", dim, dim, 1, 1, &rb->sc));
  return 0;
}
int fold_lstm(H* h, const std::string& p, int dim, int layers, LstmW* lw) {
  lw->H = dim; lw->layers = layers;
  for (int l = 0; l < layers; ++l) {
    const std::string s = std::to_string(l);
    lw->wih[l] = Wp(h, p + ".lstm.weight_ih_l" + s);
    lw->whh[l] = Wp(h, p + ".lstm.weight_hh_l" + s);
    {
      float* wt = nullptr;
      TRY(dalloc(h, &wt, (size_t)4 * dim * dim));
      TRY(conv_w_transpose_launch(lw->wih[l], wt, 4 * dim, dim, 1, 1, 0, 0, 0));
      lw->wih_t[l] = wt;
      TRY(prep_codec_tc(h, wt, dim, 1, 4 * dim, &lw->tcw[l]));
    }
    float* b = nullptr;
    TRY(dalloc(h, &b, (size_t)4 * dim));
    TRY(add_vec_launch(Wp(h, p + ".lstm.bias_ih_l" + s), Wp(h, p + ".lstm.bias_hh_l" + s), b, 4 * dim, 0));
    lw->bias[l] = b;
  }
  return 0;
}

int fold_codec(H* h) {
  const LadiffConfig& c = h->cfg;
  const int nf = c.n_filters, dim = c.rep_dims, nr = c.n_enc_ratios;
  {
    int i = 0, mult = 1;
    TRY(fold_wn_conv(h, M("encoder", i++), nf, 1, 7, 1, &h->enc.first));
    for (int r = nr - 1; r >= 0; --r) {
      ResBlockW rb; TRY(fold_resblock(h, M("encoder", i++), mult * nf, &rb)); h->enc.rb.push_back(rb);
      i++;
      ConvW d; TRY(fold_wn_conv(h, M("encoder", i++), mult * nf * 2, mult * nf, 2 * c.enc_ratios[r], c.enc_ratios[r], &d));
      h->enc.down.push_back(d);
      mult *= 2;
    }
    if (c.lstm_layers) TRY(fold_lstm(h, M("encoder", i++), mult * nf, c.lstm_layers, &h->enc.lstm));
    i++;
    TRY(fold_wn_conv(h, M("encoder", i), dim, mult * nf, 7, 1, &h->enc.last));
    // the encoder output is quantised (argmin over codewords): its tensor-core convs use the extra accumulator chains (codec_tc.cu)
    h->enc.first.precise = h->enc.last.precise = h->enc.lstm.precise = 1;
    for (auto& d : h->enc.down) d.precise = 1;
    for (auto& rb : h->enc.rb) rb.c1.precise = rb.c2.precise = rb.sc.precise = 1;
  }
  {
    int i = 0, mult = 1 << nr;
    TRY(fold_wn_conv(h, M("decoder", i++), mult * nf, dim, 7, 1, &h->dec.first));
    if (c.lstm_layers) TRY(fold_lstm(h, M("decoder", i++), mult * nf, c.lstm_layers, &h->dec.lstm));
    for (int r = 0; r < nr; ++r) {
      i++;
      ConvTrW u; TRY(fold_wn_convtr(h, M("decoder", i++), mult * nf, mult * nf / 2, c.enc_ratios[r], &u)); h->dec.up.push_back(u);
      ResBlockW rb; TRY(fold_resblock(h, M("decoder", i++), mult * nf / 2, &rb)); h->dec.rb.push_back(rb);
      mult /= 2;
    }
    i++;
    TRY(fold_wn_conv(h, M("decoder", i), 1, nf, 7, 1, &h->dec.last));
  }
  if (c.quantization) {
    TRY(dalloc(h, &h->embed, (size_t)c.n_q * kBins * dim));
    TRY(dalloc(h, &h->embed_sq, (size_t)c.n_q * kBins));
    for (int q = 0; q < c.n_q; ++q) {
      const std::string p = "quantizer.vq.layers." + std::to_string(q) + "._codebook";
      LADIFF_CUDA_OK(cudaMemcpy(h->embed + (size_t)q * kBins * dim, Wp(h, p + ".embed"), sizeof(float) * kBins * dim,
                                cudaMemcpyDeviceToDevice));
      float inited = 0.f;
      LADIFF_CUDA_OK(cudaMemcpy(&inited, Wp(h, p + ".inited"), sizeof(float), cudaMemcpyDeviceToHost));
      LADIFF_REQUIRE(inited != 0.f, LADIFF_ERR_UNSUPPORTED,
                     "%s.inited == 0: the reference would run k-means on the input batch (core_vq.py:209); not on the sampling path",
                     p.c_str());
    }
    TRY(rowsq_launch(h->embed, h->embed_sq, c.n_q * kBins, dim, 0));
    TRY(dalloc(h, &h->embed_t, (size_t)c.n_q * kBins * dim));
    TRY(rvq_transpose_launch(h->embed, h->embed_t, c.n_q, kBins, dim, 0));
  }
  return 0;
}

int pack_unet_conv(H* h, const std::string& wname, const std::string& bname, int cout, int cin, int k, int kind, bool standardize,
                   PackedConv* pc) {
  const float* w = Wp(h, wname);
  LADIFF_REQUIRE(w != nullptr, LADIFF_ERR_KEY, "missing %s", wname.c_str());
  LADIFF_REQUIRE(cin % 64 == 0, LADIFF_ERR_UNSUPPORTED, "%s: Cin=%d is not a multiple of 64", wname.c_str(), cin);
  pc->Cin = cin; pc->K = k; pc->kind = kind;
  pc->bias = bname.empty() ? nullptr : Wp(h, bname);
  if (kind == CK_UP) {
    pc->CoutV = 2 * cout; pc->Ktot = 3 * cin;
    TRY(dalloc(h, &pc->w, (size_t)pc->CoutV * pc->Ktot));
    TRY(pack_up_launch(w, pc->w, cout, cin, 0));
    float* b2 = nullptr;
    TRY(dalloc(h, &b2, (size_t)2 * cout));
    TRY(dup_bias_launch(pc->bias, b2, cout, 0));
    pc->bias = b2;
  } else {
    pc->CoutV = cout; pc->Ktot = k * cin;
    TRY(dalloc(h, &pc->w, (size_t)cout * pc->Ktot));
    TRY(pack_conv_launch(w, pc->w, cout, cin, k, standardize ? 1 : 0, 0));
  }
  LADIFF_REQUIRE(pc->CoutV % 128 == 0, LADIFF_ERR_UNSUPPORTED, "%s: Cout=%d is not a multiple of 128", wname.c_str(), pc->CoutV);
  TRY(tc_make_tmap_w(&pc->tmW, pc->w, pc->CoutV, pc->Ktot));
  return 0;
}

// block1.proj (weight-standardised k3) with res_conv (1x1, when Cin != Cout) fused as extra output rows [cout, 2cout):
// both read the same input tile, so the 1x1 shares block1's activation loads and costs no launch of its own
int pack_block1_fused(H* h, const std::string& p, int cin, int cout, PackedConv* pc) {
  const float* w1 = Wp(h, p + ".block1.proj.weight");
  const float* wr = Wp(h, p + ".res_conv.weight");
  LADIFF_REQUIRE(w1 && wr, LADIFF_ERR_KEY, "missing %s block1/res_conv weights", p.c_str());
  LADIFF_REQUIRE(cin % 64 == 0 && cout % 128 == 0, LADIFF_ERR_UNSUPPORTED, "%s: Cin=%d Cout=%d", p.c_str(), cin, cout);
  pc->Cin = cin; pc->K = 3; pc->kind = CK_PLAIN; pc->CoutV = 2 * cout; pc->Ktot = 3 * cin; pc->split_m = cout;
  TRY(dalloc(h, &pc->w, (size_t)pc->CoutV * pc->Ktot));
  LADIFF_CUDA_OK(cudaMemset(pc->w, 0, sizeof(h16) * (size_t)pc->CoutV * pc->Ktot));
  TRY(pack_conv_launch(w1, pc->w, cout, cin, 3, 1, 0));
  TRY(pack_conv_launch(wr, pc->w + (size_t)cout * pc->Ktot, cout, cin, 1, 0, 0, pc->Ktot));
  float* b2 = nullptr;
  TRY(dalloc(h, &b2, (size_t)2 * cout));
  LADIFF_CUDA_OK(cudaMemcpy(b2, Wp(h, p + ".block1.proj.bias"), sizeof(float) * cout, cudaMemcpyDeviceToDevice));
  LADIFF_CUDA_OK(cudaMemcpy(b2 + cout, Wp(h, p + ".res_conv.bias"), sizeof(float) * cout, cudaMemcpyDeviceToDevice));
  pc->bias = b2;
  TRY(tc_make_tmap_w(&pc->tmW, pc->w, pc->CoutV, pc->Ktot));
  return 0;
}

int fold_resnet(H* h, const std::string& p, int cin, int cout, long long* film_cursor, ResnetW* r) {
  r->Cin = cin; r->Cout = cout;
  r->has_res = cin != cout;
  if (r->has_res) TRY(pack_block1_fused(h, p, cin, cout, &r->c1));
  else TRY(pack_unet_conv(h, p + ".block1.proj.weight", p + ".block1.proj.bias", cout, cin, 3, CK_PLAIN, true, &r->c1));
  TRY(pack_unet_conv(h, p + ".block2.proj.weight", p + ".block2.proj.bias", cout, cout, 3, CK_PLAIN, true, &r->c2));
  r->g1 = Wp(h, p + ".block1.norm.weight"); r->b1 = Wp(h, p + ".block1.norm.bias");
  r->g2 = Wp(h, p + ".block2.norm.weight"); r->b2 = Wp(h, p + ".block2.norm.bias");
  r->film_off = *film_cursor;
  *film_cursor += 2 * cout;
  return 0;
}
int fold_attn(H* h, const std::string& p, int C, bool linear, AttnW* a) {
  a->C = C;
  TRY(pack_unet_conv(h, p + ".fn.fn.to_qkv.weight", "", 384, C, 1, CK_PLAIN, false, &a->qkv));
  if (linear) {
    TRY(pack_unet_conv(h, p + ".fn.fn.to_out.0.weight", p + ".fn.fn.to_out.0.bias", C, 128, 1, CK_PLAIN, false, &a->out));
    a->out_g = Wp(h, p + ".fn.fn.to_out.1.g");
  } else {
    TRY(pack_unet_conv(h, p + ".fn.fn.to_out.weight", p + ".fn.fn.to_out.bias", C, 128, 1, CK_PLAIN, false, &a->out));
    a->out_g = nullptr;
  }
  a->norm_g = Wp(h, p + ".fn.norm.g");
  return 0;
}

struct FilmJob { std::string prefix; long long off; int cout; };

int fold_unet(H* h) {
  const LadiffConfig& c = h->cfg;
  UNetW& u = h->un;
  const std::string pre = "diff_model";
  const int dim = c.diff_dims, inp = c.rep_dims, td = 4 * dim;
  LADIFF_REQUIRE(dim % 128 == 0 && inp == 128 && c.cond_channels == 128, LADIFF_ERR_UNSUPPORTED,
                 "UNet kernels need diff_dims %% 128 == 0 and rep_dims == cond_channels == 128 (got %d, %d, %d)", dim, inp,
                 c.cond_channels);
  u.dims[0] = dim;
  for (int i = 0; i < 5; ++i) u.dims[i + 1] = dim * kMults[i];
  long long fc = 0;
  std::vector<FilmJob> jobs;
  auto RES = [&](const std::string& p, int cin, int cout, ResnetW* r) -> int {
    jobs.push_back(FilmJob{p, fc, cout});
    return fold_resnet(h, p, cin, cout, &fc, r);
  };
  TRY(pack_unet_conv(h, pre + ".init_conv.weight", pre + ".init_conv.bias", dim, inp + c.cond_channels, 7, CK_PLAIN, false, &u.init));
  for (int i = 0; i < 5; ++i) {
    const std::string p = pre + ".downs." + std::to_string(i);
    TRY(RES(p + ".0", u.dims[i], u.dims[i], &u.d[i][0]));
    TRY(RES(p + ".1", u.dims[i], u.dims[i], &u.d[i][1]));
    TRY(fold_attn(h, p + ".2", u.dims[i], true, &u.da[i]));
    TRY(pack_unet_conv(h, p + ".3.weight", p + ".3.bias", u.dims[i + 1], u.dims[i], i < 4 ? 4 : 3, i < 4 ? CK_DOWN : CK_PLAIN, false,
                       &u.down[i]));
  }
  TRY(RES(pre + ".mid_block1", u.dims[5], u.dims[5], &u.mid1));
  TRY(fold_attn(h, pre + ".mid_attn", u.dims[5], false, &u.mida));
  TRY(RES(pre + ".mid_block2", u.dims[5], u.dims[5], &u.mid2));
  for (int j = 0; j < 5; ++j) {
    const int i = 4 - j, di = u.dims[i], dout = u.dims[i + 1];
    const std::string p = pre + ".ups." + std::to_string(j);
    TRY(RES(p + ".0", dout + di, dout, &u.u[j][0]));
    TRY(RES(p + ".1", dout + di, dout, &u.u[j][1]));
    TRY(fold_attn(h, p + ".2", dout, true, &u.ua[j]));
    const std::string q = p + (j < 4 ? ".3.1" : ".3");
    TRY(pack_unet_conv(h, q + ".weight", q + ".bias", di, dout, 3, j < 4 ? CK_UP : CK_PLAIN, false, &u.up[j]));
  }
  TRY(RES(pre + ".final_res_block", 2 * dim, dim, &u.fin));
  TRY(pack_unet_conv(h, pre + ".final_conv.weight", pre + ".final_conv.bias", inp, dim, 1, CK_PLAIN, false, &u.finalc));
  // FiLM table: time_mlp (unet.py:327-332) and every block's mlp (unet.py:162-165,183-186) depend only on t
  u.film_stride = fc;
  TRY(dalloc(h, &u.film, (size_t)kTimesteps * fc));
  float *emb = nullptr, *t1 = nullptr, *t2 = nullptr;
  LADIFF_CUDA_OK(cudaMalloc(&emb, sizeof(float) * kTimesteps * dim));
  LADIFF_CUDA_OK(cudaMalloc(&t1, sizeof(float) * kTimesteps * td));
  LADIFF_CUDA_OK(cudaMalloc(&t2, sizeof(float) * kTimesteps * td));
  int rc = sinusoid_launch(emb, kTimesteps, dim, 0);
  if (!rc) rc = linear_f32_launch(emb, dim, Wp(h, pre + ".time_mlp.1.weight"), Wp(h, pre + ".time_mlp.1.bias"), t1, td, kTimesteps, td,
                                  dim, 0, 1, 0);
  if (!rc) rc = linear_f32_launch(t1, td, Wp(h, pre + ".time_mlp.3.weight"), Wp(h, pre + ".time_mlp.3.bias"), t2, td, kTimesteps, td, td,
                                  0, 0, 0);
  for (size_t k = 0; !rc && k < jobs.size(); ++k)
    rc = linear_f32_launch(t2, td, Wp(h, jobs[k].prefix + ".mlp.1.weight"), Wp(h, jobs[k].prefix + ".mlp.1.bias"),
                           u.film + jobs[k].off, (int)fc, kTimesteps, 2 * jobs[k].cout, td, 1, 0, 0);
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(emb); cudaFree(t1); cudaFree(t2);
  if (rc) return rc;
  LADIFF_CUDA_OK(e);
  u.tb.sqrt_recip_ac = Wp(h, "diffusion.sqrt_recip_alphas_cumprod");
  u.tb.sqrt_recipm1_ac = Wp(h, "diffusion.sqrt_recipm1_alphas_cumprod");
  u.tb.coef1 = Wp(h, "diffusion.posterior_mean_coef1");
  u.tb.coef2 = Wp(h, "diffusion.posterior_mean_coef2");
  u.tb.logvar = Wp(h, "diffusion.posterior_log_variance_clipped");
  {
    auto host = [&](const char* name, std::vector<float>& v) -> int {
      v.resize(kTimesteps);
      LADIFF_CUDA_OK(cudaMemcpy(v.data(), Wp(h, std::string("diffusion.") + name), sizeof(float) * kTimesteps, cudaMemcpyDeviceToHost));
      return 0;
    };
    TRY(host("sqrt_recip_alphas_cumprod", u.h_recip));
    TRY(host("sqrt_recipm1_alphas_cumprod", u.h_recipm1));
    TRY(host("posterior_mean_coef1", u.h_c1));
    TRY(host("posterior_mean_coef2", u.h_c2));
    TRY(host("posterior_log_variance_clipped", u.h_logvar));
    TRY(host("alphas_cumprod", u.h_ac));
    u.d_sqrt_ac = Wp(h, "diffusion.sqrt_alphas_cumprod");
    u.d_sqrt_1m_ac = Wp(h, "diffusion.sqrt_one_minus_alphas_cumprod");
    u.d_p2w = Wp(h, "diffusion.p2_loss_weight");
  }
  for (int j = 0; j < c.n_upsampling_ratios; ++j) {   // cond upsamplers: plain ConvTranspose1d (no weight-norm)
    const std::string p = pre + ".upsampling_layers." + std::to_string(j) + ".convtr.convtr";
    const int s = c.upsampling_ratios[j], cc = c.cond_channels;
    ConvTrW t;
    float* w2 = nullptr;
    TRY(dalloc(h, &w2, (size_t)cc * cc * 2 * s));
    TRY(convtr_pack_launch(Wp(h, p + ".weight"), w2, cc, cc, s, 0));
    t.w2 = w2; t.bias = Wp(h, p + ".bias"); t.Cin = cc; t.Cout = cc; t.s = s;
    {
      float* wt = nullptr;
      TRY(dalloc(h, &wt, (size_t)cc * cc * 2 * s));
      TRY(conv_w_transpose_launch(w2, wt, s * cc, cc, 2, 1, s, cc, 0));
      t.wt = wt;
      TRY(prep_codec_tc(h, wt, cc, 2, s * cc, &t.tcw));
    }
    u.cond_up.push_back(t);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ codec stages
struct Pool {   // rotating activation buffers of equal size
  std::vector<float*> free_;
  float* get() { float* p = free_.back(); free_.pop_back(); return p; }
  void put(float* p) { free_.push_back(p); }
};

int conv_out_len(int Lin, int K, int stride) {   // conv.py:56-63 with padding_total = (K-1) - (stride-1)
  const int pt = (K - 1) - (stride - 1);
  return (Lin - K + pt + stride - 1) / stride + 1;
}

// LADIFF_CODEC_PROF=1: every codec launch is bracketed by CUDA events on its stream and printed (serialises the stage; debugging aid)
struct CodecProf {
  cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr; char label[160]; bool on;
  CodecProf(cudaStream_t s, const char* fmt, int a, int b, int c, int d, int e) : st(s) {
    static const bool env = getenv("LADIFF_CODEC_PROF") != nullptr;
    on = env;
    if (!on) return;
    snprintf(label, sizeof(label), fmt, a, b, c, d, e);
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
  }
  ~CodecProf() {
    if (!on) return;
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[codec_prof] %8.1f us  %s\n", ms * 1e3f, label);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
};

int run_conv(H* h, const ConvW& cw, const float* x, int Lin, float* y, int act_in, const float* res, int B, cudaStream_t st) {
  CodecProf cp(st, "conv Cin=%d Cout=%d K=%d stride=%d Lin=%d", cw.Cin, cw.Cout, cw.K, cw.stride, Lin);
  ConvF32Args a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.Cin = cw.Cin; a.Lin = Lin; a.w = cw.w; a.wt = cw.wt; a.bias = cw.bias; a.y = y; a.CoutV = cw.Cout;
  a.LoutV = conv_out_len(Lin, cw.K, cw.stride);
  a.K = cw.K; a.stride = cw.stride; a.padL = (cw.K - 1) - (cw.stride - 1); a.pad_reflect = 1; a.act_in = act_in; a.res = res;
  a.tcw = cw.tcw.valid ? &cw.tcw : nullptr; a.tc_precise = cw.precise;
  h->launches++;
  return conv1d_f32_launch(a, B, st);
}
// SConvTranspose1d (conv.py:252-274): y [B][Cout][Lin*s]; causal -> trim right only, else left = ceil((k-s)/2)
int run_convtr(H* h, const ConvTrW& cw, const float* x, int Lin, float* y, int act_in, bool causal, int B, cudaStream_t st) {
  CodecProf cp(st, "convtr Cin=%d Cout=%d s=%d Lin=%d causal=%d", cw.Cin, cw.Cout, cw.s, Lin, causal ? 1 : 0);
  ConvF32Args a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.Cin = cw.Cin; a.Lin = Lin; a.w = cw.w2; a.wt = cw.wt; a.bias = cw.bias; a.y = y; a.CoutV = cw.s * cw.Cout; a.LoutV = Lin + 1;
  a.K = 2; a.stride = 1; a.padL = 1; a.pad_reflect = 0; a.act_in = act_in; a.res = nullptr;
  a.il_s = cw.s; a.il_cout = cw.Cout; a.il_lout = Lin * cw.s;
  const int total = cw.s;                    // k - s
  a.il_trim = causal ? 0 : total - total / 2;
  a.tcw = cw.tcw.valid ? &cw.tcw : nullptr;
  h->launches++;
  return conv1d_f32_launch(a, B, st);
}
int run_resblock(H* h, const ResBlockW& rb, float*& x, int L, Pool& pool, int B, cudaStream_t st) {
  if (rb.c1.wt && rb.sc.wt && rb.c2.wt) {      // one fused kernel where the channel count has one (C = 32, 64)
    float* yf = pool.get();
    int rc;
    {
      CodecProf cp(st, "fused resblock C=%d L=%d B=%d%c%c", rb.sc.Cout, L, B, ' ', ' ');
      rc = seanet_resblock_launch(x, yf, rb.c1.wt, rb.c1.bias, rb.sc.wt, rb.sc.bias, rb.c2.wt, rb.c2.bias, rb.sc.Cout, B, L, st);
    }
    if (rc == 0) { h->launches++; pool.put(x); x = yf; return 0; }
    pool.put(yf);
    if (rc < 0) return rc;
  }
  float* hbuf = pool.get(); float* s = pool.get(); float* y = pool.get();
  TRY(run_conv(h, rb.c1, x, L, hbuf, 1, nullptr, B, st));
  TRY(run_conv(h, rb.sc, x, L, s, 0, nullptr, B, st));
  TRY(run_conv(h, rb.c2, hbuf, L, y, 1, s, B, st));
  pool.put(hbuf); pool.put(s); pool.put(x);
  x = y;
  return 0;
}
// SLSTM (lstm.py:22-28): y = lstm(x) + x on [B][H][T]
int run_lstm(H* h, const LstmW& lw, float*& x, int T, Pool& pool, float* hbuf, float* cbuf, int B, cudaStream_t st) {
  float* in = x;
  float* pre = pool.get();
  float* y = nullptr;
  for (int l = 0; l < lw.layers; ++l) {
    ConvW ip; ip.w = lw.wih[l]; ip.wt = lw.wih_t[l]; ip.bias = lw.bias[l]; ip.Cout = 4 * lw.H; ip.Cin = lw.H; ip.K = 1; ip.stride = 1; ip.tcw = lw.tcw[l]; ip.precise = lw.precise;
    TRY(run_conv(h, ip, in, T, pre, 0, nullptr, B, st));
    y = pool.get();
    const float* skip = (l == lw.layers - 1) ? x : nullptr;
    CodecProf cp(st, "lstm recurrence H=%d T=%d layer=%d B=%d%c", lw.H, T, l, B, ' ');
    if (lw.H == 64 || lw.H == 128) {
      TRY(lstm_seq_launch(pre, lw.whh[l], skip, y, B, lw.H, T, st));
      h->launches++;
    } else if (lw.H % 32 == 0 && lw.H / 4 <= tc_num_sms() && lw.H <= 1024) {
      TRY(lstm_persist_launch(pre, lw.whh[l], skip, y, hbuf, reinterpret_cast<unsigned int*>(cbuf), B, lw.H, T, st));
      h->launches++;
    } else {
      TRY(lstm_steps_launch(pre, lw.whh[l], skip, y, hbuf, cbuf, B, lw.H, T, st, &h->launches));
    }
    if (in != x) pool.put(in);
    in = y;
  }
  pool.put(pre); pool.put(x);
  x = y;
  return 0;
}

size_t codec_buf_elems(const H* h, int B, int T) {
  // the largest activation on either codec path is n_filters x T; LSTM gate pre-activations are 4H x T/hop
  const LadiffConfig& c = h->cfg;
  size_t per = (size_t)c.n_filters * T;
  const int hop = h->enc_hop;
  const size_t lstm = (size_t)4 * c.n_filters * (1 << c.n_enc_ratios) * (T / hop + 1);
  if (lstm > per) per = lstm;
  return per * B + 1024;
}
const int kPoolBufs = 6;

int setup_pool(H* h, Bump& bp, int B, int T, Pool& pool, float** hbuf, float** cbuf) {
  const size_t n = codec_buf_elems(h, B, T);
  for (int i = 0; i < kPoolBufs; ++i) pool.put(bp.get<float>(n));
  const int Hmax = h->cfg.n_filters * (1 << h->cfg.n_enc_ratios);
  *hbuf = bp.get<float>(lstm_persist_scratch_floats(B, Hmax));   // >= 2*B*Hmax, the per-step fallback's need
  *cbuf = bp.get<float>((size_t)B * Hmax + 64);
  return 0;
}

int run_encoder(H* h, const float* wav, int B, int T, float* z, Bump bp, cudaStream_t st) {
  Pool pool; float *hbuf, *cbuf;
  setup_pool(h, bp, B, T, pool, &hbuf, &cbuf);
  float* x = pool.get();
  int L = T;
  TRY(run_conv(h, h->enc.first, wav, L, x, 0, nullptr, B, st));
  for (size_t i = 0; i < h->enc.down.size(); ++i) {
    TRY(run_resblock(h, h->enc.rb[i], x, L, pool, B, st));
    float* y = pool.get();
    TRY(run_conv(h, h->enc.down[i], x, L, y, 1, nullptr, B, st));
    L = conv_out_len(L, h->enc.down[i].K, h->enc.down[i].stride);
    pool.put(x); x = y;
  }
  if (h->cfg.lstm_layers) TRY(run_lstm(h, h->enc.lstm, x, L, pool, hbuf, cbuf, B, st));
  TRY(run_conv(h, h->enc.last, x, L, z, 1, nullptr, B, st));
  return 0;
}

int run_decoder(H* h, const float* zin, int B, int L, float* wav, Bump bp, cudaStream_t st) {
  Pool pool; float *hbuf, *cbuf;
  setup_pool(h, bp, B, L * h->enc_hop, pool, &hbuf, &cbuf);
  float* x = pool.get();
  TRY(run_conv(h, h->dec.first, zin, L, x, 0, nullptr, B, st));
  if (h->cfg.lstm_layers) TRY(run_lstm(h, h->dec.lstm, x, L, pool, hbuf, cbuf, B, st));
  for (size_t i = 0; i < h->dec.up.size(); ++i) {
    float* y = pool.get();
    TRY(run_convtr(h, h->dec.up[i], x, L, y, 1, true, B, st));
    L *= h->dec.up[i].s;
    pool.put(x); x = y;
    TRY(run_resblock(h, h->dec.rb[i], x, L, pool, B, st));
  }
  TRY(run_conv(h, h->dec.last, x, L, wav, 1, nullptr, B, st));
  return 0;
}

// upsampling_layers loop (sample.py:125-128 / unet.py:412-414): cond [B][128][F] -> out [B][128][F*prod]
int run_cond_upsample(H* h, const float* cond, int B, int F, float* out, float* tmp_a, float* tmp_b, cudaStream_t st) {
  const auto& ups = h->un.cond_up;
  const float* x = cond;
  int L = F;
  for (size_t j = 0; j < ups.size(); ++j) {
    float* y = (j + 1 == ups.size()) ? out : ((j & 1) ? tmp_b : tmp_a);
    TRY(run_convtr(h, ups[j], x, L, y, 0, false, B, st));
    L *= ups[j].s;
    x = y;
  }
  if (ups.empty()) LADIFF_CUDA_OK(cudaMemcpyAsync(out, cond, sizeof(float) * B * 128 * F, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------ UNet plan
void carve_unet(const H* h, Bump& bp, int B, int L, UnetBufs* u) {
  const int* d = h->un.dims;
  const size_t BL = (size_t)B * L;
  u->xin = bp.get<h16>(BL * 256);
  u->FC = bp.get<h16>(BL * 2 * d[0]);
  for (int i = 0; i < 5; ++i) {
    const size_t n = (size_t)B * (L >> i) * (d[i + 1] + d[i]);
    u->CA[i] = bp.get<h16>(n);
    u->CB[i] = bp.get<h16>(n);
  }
  u->X[0] = nullptr;
  for (int i = 1; i <= 5; ++i) u->X[i] = bp.get<h16>((size_t)B * (L >> (i < 5 ? i : 4)) * d[i]);
  size_t tmax = 0;
  for (int i = 0; i < 5; ++i) { const size_t n = (size_t)B * (L >> i) * d[i + 1]; if (n > tmax) tmax = n; }
  u->tY = bp.get<h16>(tmax); u->tH = bp.get<h16>(tmax); u->tO = bp.get<h16>(tmax); u->tR = bp.get<h16>(tmax);
  u->tA = bp.get<h16>(tmax);
  u->qkv = bp.get<h16>(BL * 384);
  u->ao = bp.get<h16>(BL * 128);
  u->eps = bp.get<float>(BL * 128);
  u->ctx = bp.get<float>((size_t)B * 4096);
  u->la_part = bp.get<float>(linattn_part_floats(B, L));
  u->la_cnt = bp.get<int>((size_t)4 * B);
  u->stats_slots = (L / 16 + 4) * 32 * TC_STAT_PARTS;
  u->stats = bp.get<float2>((size_t)B * u->stats_slots);
  u->t_dev = bp.get<int>(B);
  u->inv_scale = bp.get<float>(B);
  u->condup = bp.get<float>(BL * 128);
  u->condtmp = bp.get<float>(BL * 128);
}

ClView view(h16* p, int L, int pitch, int C, int ch0 = 0) {
  ClView v; v.p = p + ch0; v.bstride = (long long)L * pitch; v.pitch = pitch; v.C = C; return v;
}

// Timing ablations (ladiff_set_skip_ops / LADIFF_SKIP_OPS; results are garbage while a bit is set): 1 GroupNorm-apply, 2 LayerNorm,
// 4 attention cores, 8 1x1 convs, 16 all other convs.  bench.py uses it to attribute the CUDA-graph replay's time to kernel classes.
static int g_skip_env() {
  static const int m = getenv("LADIFF_SKIP_OPS") ? atoi(getenv("LADIFF_SKIP_OPS")) : 0;
  return m;
}
struct PlanBuilder {
  H* h; Plan* pl; int B;
  bool tune_stream_ok; cudaStream_t tune_stream; cudaEvent_t tune_ev[2];
  void label(const char* fmt, int a, int b, int c, int d) {
    char buf[120];
    snprintf(buf, sizeof(buf), fmt, a, b, c, d);
    pl->op_label.resize(pl->ops.size());
    pl->op_label.back() = buf;
  }
  // conv: `in` has Lin rows; writes Lout rows into `out` (h16) or out32 (fp32, contiguous [B][Lout][CoutV])
  int conv(const PackedConv& pc, ClView in, int Lin, ClView out, float* out32, bool want_stats, int* n_ntiles, ClView res,
           ClView out2 = ClView{nullptr, 0, 0, 0}) {
    LADIFF_REQUIRE(in.C == pc.Cin, LADIFF_ERR_ARG, "plan: conv input has %d channels, weights expect %d", in.C, pc.Cin);
    TcConvDesc d;
    memset(&d, 0, sizeof(d));
    d.kind = pc.kind; d.Cin = pc.Cin; d.K = pc.K; d.CoutV = pc.CoutV; d.w = pc.w; d.tmW = &pc.tmW; d.Ktot = pc.Ktot; d.bias = pc.bias;
    d.x = in.p; d.x_bstride = in.bstride; d.x_pitch = in.pitch; d.Lin = Lin;
    d.out = out.p; d.out_bstride = out.bstride; d.out_pitch = out.pitch; d.out32 = out32;
    d.stats = want_stats ? pl->bufs.stats : nullptr;
    if (res.p) { d.res = res.p; d.res_bstride = res.bstride; d.res_pitch = res.pitch; }
    if (pc.split_m) {
      LADIFF_REQUIRE(out2.p != nullptr, LADIFF_ERR_ARG, "plan: fused res_conv needs a second output");
      d.split_m = pc.split_m; d.out2 = out2.p; d.out2_bstride = out2.bstride; d.out2_pitch = out2.pitch;
    }
    d.B = B;
    TcConvParams ps, pu;
    TcRefView rv;
    d.tap_share = 1;
    TRY(tc_conv_plan(d, &ps, &rv));
    char tkey[160];
    snprintf(tkey, sizeof(tkey), "%d:%d:%d:%d:%d:%d:%d:%d:%d:%d", pc.kind, pc.Cin, pc.K, pc.CoutV, pc.split_m, Lin, B, want_stats ? 1 : 0,
             out32 ? 1 : 0, res.p ? 1 : 0);
    auto tuned_it = h->tuned.find(tkey);
    if (tuned_it != h->tuned.end() && h->conv_impl == 0) {        // an earlier plan of this handle already timed this conv
      const std::vector<int>& c = tuned_it->second;
      TcConvDesc dc = d;
      dc.want_nt = c[0]; dc.want_nclip = c[1]; dc.want_two_per_sm = c[2]; dc.want_transposed = c[3]; dc.want_pair = c[4];
      TcConvParams pc2; TcRefView rv2;
      if (tc_conv_plan(dc, &pc2, &rv2) == 0) { ps = pc2; rv = rv2; }
    } else
    if (tune_stream_ok && h->conv_impl == 0) {
      // Autotune the tile shape on the real buffers: wave quantisation over 148 SMs, operand traffic and epilogue cost all
      // depend on it and none is monotonic in the tile width.  A candidate replaces the cost model's pick only if it is >3 %
      // faster (keeps the choice stable against timing noise).  The conv's inputs hold arbitrary data here; its outputs are
      // scratch that the real evaluation overwrites.
      auto time_one = [&](const TcConvParams& cand, float* ms) -> int {
        TRY(tc_conv_launch(cand, tune_stream));
        LADIFF_CUDA_OK(cudaEventRecord(tune_ev[0], tune_stream));
        for (int r = 0; r < 3; ++r) TRY(tc_conv_launch(cand, tune_stream));
        LADIFF_CUDA_OK(cudaEventRecord(tune_ev[1], tune_stream));
        LADIFF_CUDA_OK(cudaEventSynchronize(tune_ev[1]));
        LADIFF_CUDA_OK(cudaEventElapsedTime(ms, tune_ev[0], tune_ev[1]));
        return 0;
      };
      float best_ms = 0.f;
      TRY(time_one(ps, &best_ms));
      const float base_ms = best_ms;
      TcConvParams best = ps;
      TcRefView best_rv = rv;
      const bool multi = ps.NCLIP > 1 || ps.Lout + 8 <= 120;
      static const bool no_t = getenv("LADIFF_NO_TRANSPOSED") != nullptr;
      static const bool no_two = getenv("LADIFF_NO_TWO_PER_SM") != nullptr;
      for (int c = -5; c < 12; ++c) {
        TcConvDesc dc = d;
        if (c < -2) { if (no_two || multi) continue; dc.want_two_per_sm = 1; dc.want_nt = 128 - 16 * (-3 - c); }   // two CTAs per SM: NT = 128, 112, 96
        else if (c < 0) { if (no_t || ps.Lout < 128) continue; dc.want_transposed = -c; }    // positions-on-M kernel: one CTA / CTA pair per tile
        else if (multi) { dc.want_nclip = c + 1; if (c >= 3) break; }
        else { dc.want_nt = 256 - 16 * c; if (dc.want_nt < 96) break; }
        TcConvParams pc2;
        TcRefView rv2;
        if (tc_conv_plan(dc, &pc2, &rv2) != 0) continue;                     // shape does not fit (smem, halo): skip
        if (!pc2.transposed && pc2.minb == ps.minb && pc2.NT == ps.NT && pc2.NCLIP == ps.NCLIP) continue;
        if (want_stats && pc2.n_ptiles * pc2.stat_parts * pc2.stat_slots > pl->bufs.stats_slots) continue;
        float ms = 0.f;
        TRY(time_one(pc2, &ms));
        static const int force_t = getenv("LADIFF_FORCE_TRANSPOSED") ? atoi(getenv("LADIFF_FORCE_TRANSPOSED")) : 0;     // experiment knob
        if (force_t && pc2.transposed == force_t) { best_ms = 0.f; best = pc2; best_rv = rv2; continue; }
        if (ms < best_ms && ms < 0.97f * base_ms) { best_ms = ms; best = pc2; best_rv = rv2; }
      }
      // opt-in (LADIFF_TRY_PAIR=1): measured within +-2 % of the single-CTA form on every conv shape of config 2
      // (profiles/r2c/conv_sweep_with_multicast_pairs.txt) — the main loop is MMA-issue-bound, not L2->SM-bound
      static const bool try_pair = getenv("LADIFF_TRY_PAIR") != nullptr || getenv("LADIFF_FORCE_PAIR") != nullptr;
      if (try_pair && !best.transposed && best.minb == 1 && best.n_ntiles >= 2) {
        // the chosen shape as CTA pairs along N that share every weight tile through TMA multicast (half the weight bytes per SM)
        TcConvDesc dc = d;
        dc.want_nt = best.NCLIP == 1 ? best.NT : 0; dc.want_nclip = best.NCLIP > 1 ? best.NCLIP : 0; dc.want_pair = 1;
        TcConvParams pc2; TcRefView rv2;
        if (tc_conv_plan(dc, &pc2, &rv2) == 0 && pc2.NT == best.NT && pc2.NCLIP == best.NCLIP) {
          float ms = 0.f;
          TRY(time_one(pc2, &ms));
          static const bool force_pair = getenv("LADIFF_FORCE_PAIR") != nullptr;
          if (force_pair || ms < 0.97f * best_ms) { best_ms = ms; best = pc2; best_rv = rv2; }
        }
      }
      ps = best; rv = best_rv;
      h->tuned[tkey] = std::vector<int>{ps.transposed ? 0 : (ps.NCLIP == 1 ? ps.NT : 0), ps.NCLIP > 1 ? ps.NCLIP : 0, ps.minb == 2 ? 1 : 0, ps.transposed, ps.pair};
    }
    d.tap_share = 0;
    d.want_nt = (ps.NCLIP == 1 && !ps.transposed) ? ps.NT : 0; d.want_nclip = ps.NCLIP > 1 ? ps.NCLIP : 0;
    d.want_two_per_sm = 0;
    if (ps.transposed) {          // the per-tap check variant must produce the same GroupNorm-partial layout: reuse the chosen plan
      pu = ps;
    } else
    if (tc_conv_plan(d, &pu, nullptr) != 0) { d.want_nt = 0; d.want_nclip = 0; TRY(tc_conv_plan(d, &pu, nullptr)); }
    LADIFF_REQUIRE((ps.n_ptiles == pu.n_ptiles && ps.stat_parts == pu.stat_parts) || !want_stats, LADIFF_ERR_ARG, "plan: tap-shared and per-tap tilings disagree");
    if (want_stats)
      LADIFF_REQUIRE(ps.n_ptiles * ps.stat_parts * ps.stat_slots <= pl->bufs.stats_slots, LADIFF_ERR_WORKSPACE, "plan: stats buffer too small");
    if (n_ntiles) *n_ntiles = ps.n_ptiles * ps.stat_parts;
    H* hh = h;
    const int skip_bit = pc.K == 1 ? 8 : 16;
    pl->ops.push_back([hh, ps, pu, rv, skip_bit](cudaStream_t st) {
      if ((hh->skip_ops | g_skip_env()) & skip_bit) return 0;
      if (hh->conv_impl == 1) return tc_conv_ref_launch(ps, rv, st);
      return tc_conv_launch(hh->conv_impl == 2 ? pu : ps, st);
    });
    pl->op_flops.resize(pl->ops.size(), 0.0);
    pl->op_flops.back() = pc.split_m ? 2.0 * ((double)pc.split_m * pc.Ktot + (double)(pc.CoutV - pc.split_m) * pc.Cin) * ps.Lout * B
                                     : 2.0 * pc.CoutV * (double)pc.Ktot * ps.Lout * B;
    pl->op_label.resize(pl->ops.size());
    {
      char buf[200];
      snprintf(buf, sizeof(buf), "conv kind=%d Cout=%d Cin=%d k=%d Lout=%d NT=%d nclip=%d tiles=%d S=%d taps/stage=%d stats=%d direct=%d posM=%d perSM=%d pair=%d", pc.kind,
               pc.CoutV, pc.Cin, pc.K, ps.Lout, ps.NT, ps.NCLIP, ps.transposed ? ps.n_chtiles * ps.n_ntiles : ps.MT * ps.n_ntiles, ps.S, ps.a_cap,
               want_stats ? 1 : 0, ps.direct, ps.transposed ? ps.NCH : 0, ps.transposed ? 1 : ps.minb, ps.pair);
      pl->op_label.back() = buf;
    }
    pl->launches_per_run++;
    return 0;
  }
  int gn(ClView y, int L, int n_ntiles, const float* g, const float* b, const float* film, ClView res, ClView out, bool do_tanh,
         const float* ln_g = nullptr, ClView ln_out = ClView{nullptr, 0, 0, 0}) {
    GnApplyArgs a;
    memset(&a, 0, sizeof(a));
    a.ln_g = ln_g; a.ln_out = ln_out;
    a.y = y; a.stats = pl->bufs.stats; a.n_ntiles = n_ntiles; a.gamma = g; a.beta = b; a.film = film;
    a.film_stride = h->un.film_stride; a.t_dev = pl->bufs.t_dev; a.res = res; a.out = out; a.L = L; a.do_tanh = do_tanh ? 1 : 0;
    const int BB = B;
    H* hg = h;
    pl->ops.push_back([a, BB, hg](cudaStream_t st) { return ((hg->skip_ops | g_skip_env()) & 1) ? 0 : gn_apply_launch(a, BB, st); });
    label("gn_apply C=%d L=%d film=%d res=%d", y.C, L, film ? 1 : 0, res.p ? 1 : 0);
    pl->launches_per_run++;
    return 0;
  }
  // ResnetBlock (unet.py:176-192): out = SiLU(GN(conv2(SiLU(FiLM(GN(conv1(x))))))) + res_conv(x)
  // ln_g != null: the block's output also leaves channel-LayerNorm'ed (the following attention's pre-norm) in ln_out
  int resnet(const ResnetW& r, ClView x, int L, ClView out, bool do_tanh = false, const float* ln_g = nullptr,
             ClView ln_out = ClView{nullptr, 0, 0, 0}) {
    UnetBufs& u = pl->bufs;
    ClView none; memset(&none, 0, sizeof(none));
    ClView y = view(u.tY, L, r.Cout, r.Cout), hh = view(u.tH, L, r.Cout, r.Cout);
    int nt = 0;
    ClView resv = x;
    if (r.has_res) {                  // res_conv(x) rides along as the second output of block1's conv
      resv = view(u.tR, L, r.Cout, r.Cout);
      TRY(conv(r.c1, x, L, y, nullptr, true, &nt, none, resv));
    } else {
      TRY(conv(r.c1, x, L, y, nullptr, true, &nt, none));
    }
    TRY(gn(y, L, nt, r.g1, r.b1, h->un.film + r.film_off, none, hh, false));
    TRY(conv(r.c2, hh, L, y, nullptr, true, &nt, none));
    TRY(gn(y, L, nt, r.g2, r.b2, nullptr, resv, out, do_tanh, ln_g, ln_out));
    return 0;
  }
  int layernorm(ClView x, const float* g, ClView res, ClView out, int L) {
    const int BB = B;
    H* hg = h;
    pl->ops.push_back([x, g, res, out, BB, L, hg](cudaStream_t st) { return ((hg->skip_ops | g_skip_env()) & 2) ? 0 : layernorm_cl_launch(x, g, res, out, BB, L, st); });
    label("layernorm C=%d L=%d res=%d", x.C, L, res.p ? 1 : 0, 0);
    pl->launches_per_run++;
    return 0;
  }
  // Residual(PreNorm(LinearAttention)) (unet.py:194-222) / Residual(PreNorm(Attention)) (:224-246)
  // pre_normed: `ln` (tH) already holds LN(x)*norm_g, written by the producing GroupNorm-apply
  int attention(const AttnW& a, ClView x, int L, ClView out, bool linear, bool pre_normed = false) {
    UnetBufs& u = pl->bufs;
    ClView none; memset(&none, 0, sizeof(none));
    ClView ln = view(u.tH, L, a.C, a.C), qkv = view(u.qkv, L, 384, 384), ao = view(u.ao, L, 128, 128);
    if (!pre_normed) TRY(layernorm(x, a.norm_g, none, ln, L));
    TRY(conv(a.qkv, ln, L, qkv, nullptr, false, nullptr, none));
    const int BB = B; float* ctx = u.ctx; float* lap = u.la_part; int* lac = u.la_cnt;
    // opt-in (LADIFF_ATTN_TAIL=1): parity-green but 15 % SLOWER per DDPM step at config 2 than the three kernels it replaces — one
    // 123 KB CTA per SM runs its phases (SIMT context product, weight staging, MMA, two LayerNorm sweeps) back to back with nothing
    // to overlap them
    static const bool fuse_tail = getenv("LADIFF_ATTN_TAIL") != nullptr;
    if (linear && fuse_tail && a.C % 256 == 0 && a.C <= 1024) {
      // context, then ONE kernel for ctx^T softmax(q) -> to_out 1x1 (tcgen05) -> LayerNorm -> + x
      H* hg = h;
      pl->ops.push_back([qkv, ctx, lap, lac, BB, L, hg](cudaStream_t st) { return ((hg->skip_ops | g_skip_env()) & 4) ? 0 : linattn_ctx_launch(qkv, ctx, lap, lac, BB, L, st); });
      label("linattn ctx L=%d", L, 0, 0, 0);
      const h16* wout = a.out.w; const float* bo = a.out.bias; const float* go = a.out_g; const int C = a.C;
      pl->ops.push_back([qkv, ctx, wout, bo, go, x, out, BB, L, C, hg](cudaStream_t st) {
        return ((hg->skip_ops | g_skip_env()) & 4) ? 0 : linattn_tail_launch(qkv, ctx, wout, bo, go, x, out, BB, L, C, st);
      });
      label("linattn tail (out + to_out 1x1 + LN + res) C=%d L=%d", a.C, L, 0, 0);
      pl->op_flops.resize(pl->ops.size(), 0.0);
      pl->launches_per_run += 2;
    } else if (linear) {
      H* hg = h;
      pl->ops.push_back([qkv, ctx, lap, lac, ao, BB, L, hg](cudaStream_t st) { return ((hg->skip_ops | g_skip_env()) & 4) ? 0 : linattn_launch(qkv, ctx, lap, lac, ao, BB, L, st); });
      label("linattn(ctx+out) L=%d", L, 0, 0, 0);
      pl->launches_per_run += 2;
      ClView yo = view(u.tY, L, a.C, a.C);
      TRY(conv(a.out, ao, L, yo, nullptr, false, nullptr, none));
      TRY(layernorm(yo, a.out_g, x, out, L));
    } else {
      H* hg = h;
      pl->ops.push_back([qkv, ao, BB, L, hg](cudaStream_t st) { return ((hg->skip_ops | g_skip_env()) & 4) ? 0 : fullattn_launch(qkv, ao, BB, L, st); });
      label("fullattn L=%d", L, 0, 0, 0);
      pl->launches_per_run++;
      TRY(conv(a.out, ao, L, out, nullptr, false, nullptr, x));
    }
    return 0;
  }
};

int build_plan(H* h, void* ws_unet, int B, int L, cudaStream_t st, Plan** out) {
  static const bool dbg_plan = getenv("LADIFF_DEBUG_PLAN") != nullptr;
  for (size_t i = 0; i < h->plans.size(); ++i) {
    Plan* p = h->plans[i];
    if (p->ws == ws_unet && p->B == B && p->L == L) {      // most recently used plan at the back (LRU eviction)
      h->plans.erase(h->plans.begin() + i);
      h->plans.push_back(p);
      *out = p;
      if (dbg_plan) fprintf(stderr, "[plan] hit  %p ws=%p B=%d L=%d runs=%lld\n", (void*)p, ws_unet, B, L, p->runs);
      return 0;
    }
  }
  LADIFF_REQUIRE(L % 16 == 0 && L >= 16, LADIFF_ERR_ARG, "UNet needs a latent length that is a multiple of 16 (got %d)", L);
  Plan* pl = new Plan();
  pl->ws = ws_unet; pl->B = B; pl->L = L;
  Bump bp(ws_unet);
  carve_unet(h, bp, B, L, &pl->bufs);
  UnetBufs& u = pl->bufs;
  const UNetW& w = h->un;
  const int* d = w.dims;
  PlanBuilder pb{h, pl, B, false, st, {nullptr, nullptr}};
  {
    static const bool no_tune = getenv("LADIFF_NO_AUTOTUNE") != nullptr;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (!no_tune && cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone &&
        cudaEventCreate(&pb.tune_ev[0]) == cudaSuccess && cudaEventCreate(&pb.tune_ev[1]) == cudaSuccess)
      pb.tune_stream_ok = true;
  }
  ClView none; memset(&none, 0, sizeof(none));
  int rc = 0;
  auto CHECK = [&](int r) { if (r && !rc) rc = r; };
  static const bool fuse_ln = getenv("LADIFF_NO_LN_FUSE") == nullptr;   // attention pre-norm fused into the producing GroupNorm-apply
  // init_conv on cat(cond_up, x); its output is also `r` of the final concat (unet.py:430-435,464)
  ClView xin = view(u.xin, L, 256, 256);
  ClView r0 = view(u.FC, L, 2 * d[0], d[0], d[0]);
  CHECK(pb.conv(w.init, xin, L, r0, nullptr, false, nullptr, none));
  ClView x = r0;
  for (int i = 0; i < 5 && !rc; ++i) {
    const int Li = L >> i, C = d[i], pitchC = d[i + 1] + d[i];
    ClView skip1 = view(u.CB[i], Li, pitchC, C, d[i + 1]);
    ClView skip2 = view(u.CA[i], Li, pitchC, C, d[i + 1]);
    ClView o2 = view(u.tO, Li, C, C);
    CHECK(pb.resnet(w.d[i][0], x, Li, skip1));
    CHECK(pb.resnet(w.d[i][1], skip1, Li, o2, false, fuse_ln ? w.da[i].norm_g : nullptr, view(u.tH, Li, C, C)));
    CHECK(pb.attention(w.da[i], o2, Li, skip2, true, fuse_ln));
    const int Ln = i < 4 ? Li / 2 : Li;
    ClView xn = view(u.X[i + 1], Ln, d[i + 1], d[i + 1]);
    CHECK(pb.conv(w.down[i], skip2, Li, xn, nullptr, false, nullptr, none));
    x = xn;
  }
  const int Lm = L >> 4, Cm = d[5];
  if (!rc) {
    ClView m1 = view(u.tO, Lm, Cm, Cm), m2 = view(u.tA, Lm, Cm, Cm);
    ClView m3 = view(u.CA[4], Lm, d[5] + d[4], d[5], 0);
    CHECK(pb.resnet(w.mid1, x, Lm, m1, false, fuse_ln ? w.mida.norm_g : nullptr, view(u.tH, Lm, Cm, Cm)));
    CHECK(pb.attention(w.mida, m1, Lm, m2, false, fuse_ln));
    CHECK(pb.resnet(w.mid2, m2, Lm, m3));
  }
  for (int j = 0; j < 5 && !rc; ++j) {
    const int i = 4 - j, Li = L >> i, Cout = d[i + 1], pitchC = d[i + 1] + d[i];
    ClView ca = view(u.CA[i], Li, pitchC, pitchC), cb = view(u.CB[i], Li, pitchC, pitchC);
    ClView b1 = view(u.CB[i], Li, pitchC, Cout, 0);
    ClView o2 = view(u.tO, Li, Cout, Cout), at = view(u.tA, Li, Cout, Cout);
    CHECK(pb.resnet(w.u[j][0], ca, Li, b1));
    CHECK(pb.resnet(w.u[j][1], cb, Li, o2, false, fuse_ln ? w.ua[j].norm_g : nullptr, view(u.tH, Li, Cout, Cout)));
    CHECK(pb.attention(w.ua[j], o2, Li, at, true, fuse_ln));
    if (j < 4) {
      const int pn = d[i] + d[i - 1];   // pitch of CA[i-1]; x part = channels [0, d[i])
      ClView dst = view(u.CA[i - 1], 2 * Li, pn, d[i], 0);
      dst.bstride = (long long)(2 * Li) * pn;
      CHECK(pb.conv(w.up[j], at, Li, dst, nullptr, false, nullptr, none));
    } else {
      ClView dst = view(u.FC, L, 2 * d[0], d[0], 0);
      CHECK(pb.conv(w.up[j], at, Li, dst, nullptr, false, nullptr, none));
    }
  }
  if (!rc) {
    ClView fc = view(u.FC, L, 2 * d[0], 2 * d[0]);
    ClView fo = view(u.tO, L, d[0], d[0]);
    CHECK(pb.resnet(w.fin, fc, L, fo, true));                        // + tanh (unet.py:467)
    CHECK(pb.conv(w.finalc, fo, L, none, u.eps, false, nullptr, none));
  }
  if (pb.tune_ev[0]) cudaEventDestroy(pb.tune_ev[0]);
  if (pb.tune_ev[1]) cudaEventDestroy(pb.tune_ev[1]);
  if (rc) { delete pl; return rc; }
  if (h->plans.size() >= 6) {                                // evict the least recently used plan
    if (h->last_plan == h->plans.front()) h->last_plan = nullptr;
    delete h->plans.front();
    h->plans.erase(h->plans.begin());
  }
  h->plans.push_back(pl);
  *out = pl;
  if (dbg_plan) fprintf(stderr, "[plan] built %p ws=%p B=%d L=%d (%zu cached)\n", (void*)pl, ws_unet, B, L, h->plans.size());
  return 0;
}

int run_plan(H* h, Plan* pl, cudaStream_t st) {
  pl->op_flops.resize(pl->ops.size(), 0.0);
  if (!h->profiling) {
    // The evaluation is a fixed launch sequence over fixed buffers (the step index lives in t_dev), so after one eager
    // run (which also performs the one-time cudaFuncSetAttribute calls) it is captured into a CUDA graph and replayed:
    // ~170 kernel launches per DDPM step become one graph launch.
    static const bool no_graph = getenv("LADIFF_NO_GRAPH") != nullptr || getenv("LADIFF_TC_PROF") != nullptr;
    if (no_graph || pl->runs == 0) {
      for (auto& op : pl->ops) TRY(op(st));
    } else {
      if (pl->gexec && pl->graph_impl != h->conv_impl) { cudaGraphExecDestroy(pl->gexec); pl->gexec = nullptr; }
      if (!pl->gexec) {
        cudaGraph_t g = nullptr;
        // capture on a private stream: the caller's stream may be the legacy default stream, which cannot capture
        if (!pl->cap_stream) LADIFF_CUDA_OK(cudaStreamCreateWithFlags(&pl->cap_stream, cudaStreamNonBlocking));
        LADIFF_CUDA_OK(cudaStreamBeginCapture(pl->cap_stream, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (auto& op : pl->ops) { rc = op(pl->cap_stream); if (rc) break; }
        cudaError_t e = cudaStreamEndCapture(pl->cap_stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        LADIFF_CUDA_OK(e);
        e = cudaGraphInstantiate(&pl->gexec, g, 0);
        cudaGraphDestroy(g);
        LADIFF_CUDA_OK(e);
        pl->graph_impl = h->conv_impl;
      }
      LADIFF_CUDA_OK(cudaGraphLaunch(pl->gexec, st));
    }
    pl->runs++;
  } else {   // CUDA events on the launching stream around every conv launch (bench.py roofline)
    if (pl->ev.empty()) {
      pl->ev.resize(2 * pl->ops.size() + 2);
      for (auto& e : pl->ev) LADIFF_CUDA_OK(cudaEventCreate(&e));
    }
    LADIFF_CUDA_OK(cudaEventRecord(pl->ev[2 * pl->ops.size()], st));
    for (size_t i = 0; i < pl->ops.size(); ++i) {
      const bool timed = pl->op_flops[i] > 0.0 || h->profiling == 2;
      if (timed) LADIFF_CUDA_OK(cudaEventRecord(pl->ev[2 * i], st));
      TRY(pl->ops[i](st));
      if (timed) LADIFF_CUDA_OK(cudaEventRecord(pl->ev[2 * i + 1], st));
    }
    LADIFF_CUDA_OK(cudaEventRecord(pl->ev[2 * pl->ops.size() + 1], st));
    h->last_plan = pl;
  }
  h->launches += pl->launches_per_run;
  return 0;
}

// process_cond (unet.py:407-420) into xin[:, :, 0:128]
int prepare_cond(H* h, Plan* pl, const float* cond, int B, int L, int F, cudaStream_t st) {
  UnetBufs& u = pl->bufs;
  int Lup = F;
  for (auto& c : h->un.cond_up) Lup *= c.s;
  LADIFF_REQUIRE(Lup == L, LADIFF_ERR_ARG, "cond frames %d x upsampling = %d != latent length %d", F, Lup, L);
  LADIFF_CUDA_OK(cudaMemsetAsync(u.la_cnt, 0, sizeof(int) * 4 * B, st));   // linear-attention tickets (the workspace is shared with the codec stages)
  TRY(run_cond_upsample(h, cond, B, F, u.condup, u.condtmp, (float*)u.eps, st));
  const float* inv = nullptr;
  if (h->cfg.unet_scale_cond) {
    TRY(absmax_inv_launch(u.condup, u.inv_scale, B, (long long)128 * L, 1e-20f, st));
    inv = u.inv_scale; h->launches++;
  }
  TRY(ncl_to_cl_launch(u.condup, inv, view(u.xin, L, 256, 128, 0), B, 128, L, st));
  h->launches++;
  return 0;
}

size_t unet_ws_bytes(const H* h, int B, int L) {
  Bump bp(nullptr);
  UnetBufs u;
  carve_unet(h, bp, B, L, &u);
  return bp.off + 8192;
}
size_t codec_ws_bytes(const H* h, int B, int T) {
  const size_t n = codec_buf_elems(h, B, T);
  const int Hmax = h->cfg.n_filters * (1 << h->cfg.n_enc_ratios);
  const size_t z = align_up(sizeof(float) * (size_t)B * h->cfg.rep_dims * (T / h->enc_hop + 1), 1024);   // get_cond's encoder output
  return (n * kPoolBufs + lstm_persist_scratch_floats(B, Hmax) + (size_t)B * Hmax + 64) * sizeof(float) + z + 32 * 1024;
}
// persistent region used by ladiff_synthesize: cond [B][128][F], x [B][128][L], two upsample temporaries
size_t persist_bytes(int B, int T, int L) {
  return align_up(((size_t)B * 128 * (T / 320 + 1) + (size_t)3 * B * 128 * L) * sizeof(float) + 8 * 1024, 1024);
}

int check_ws(void* ws, int64_t have, size_t need) {
  LADIFF_REQUIRE(ws != nullptr && ((uintptr_t)ws % 256) == 0, LADIFF_ERR_WORKSPACE, "workspace must be non-null and 256-byte aligned");
  LADIFF_REQUIRE((size_t)have >= need, LADIFF_ERR_WORKSPACE, "workspace too small: have %lld, need %zu", (long long)have, need);
  return 0;
}

// p_sample's coefficients at step t (ddpm_loss.py:199-206, 244-251); sigma = exp(0.5 logvar) in fp32 like the reference
StepCoef ddpm_coef(const UNetW& u, int t) {
  StepCoef c;
  c.a = u.h_recip[t]; c.b = u.h_recipm1[t]; c.k0 = u.h_c1[t]; c.k1 = u.h_c2[t];
  c.ks = t > 0 ? expf(0.5f * u.h_logvar[t]) : 0.f;
  c.mode = 0;
  return c;
}
// ddim_sample's coefficients for the pair (time, time_next) (ddpm_loss.py:289-300), in the reference's fp32 operation order
StepCoef ddim_coef(const UNetW& u, int time, int time_next, float eta) {
  StepCoef c;
  c.a = u.h_recip[time]; c.b = u.h_recipm1[time];
  if (time_next < 0) { c.k0 = 1.f; c.k1 = 0.f; c.ks = 0.f; c.mode = 2; return c; }
  const float alpha = u.h_ac[time], alpha_next = u.h_ac[time_next];
  const float sigma = eta * sqrtf((1.f - alpha / alpha_next) * (1.f - alpha_next) / (1.f - alpha));
  const float cc = sqrtf(1.f - alpha_next - sigma * sigma);
  c.k0 = sqrtf(alpha_next); c.k1 = cc; c.ks = sigma; c.mode = 1;
  return c;
}

int ddpm_run(H* h, Plan* pl, float* x, const float* cond, const float* noise, int64_t n_noise, uint64_t seed, int t_start, int n_steps,
             int B, int L, int F, cudaStream_t st) {
  UnetBufs& u = pl->bufs;
  LADIFF_REQUIRE(t_start <= kTimesteps && n_steps >= 0 && t_start - n_steps >= 0, LADIFF_ERR_ARG, "ddpm: t_start=%d n_steps=%d",
                 t_start, n_steps);
  if (getenv("LADIFF_DEBUG_PLAN")) fprintf(stderr, "[ddpm] x=%p cond=%p noise=%p n_noise=%lld t_start=%d n=%d B=%d L=%d F=%d\n", (void*)x, (void*)cond, (void*)noise, (long long)n_noise, t_start, n_steps, B, L, F);
  TRY(prepare_cond(h, pl, cond, B, L, F, st));
  ClView xs = view(u.xin, L, 256, 128, 128);
  TRY(ncl_to_cl_launch(x, nullptr, xs, B, 128, L, st));
  h->launches++;
  int64_t k = 0;
  for (int s = 0; s < n_steps; ++s) {
    const int t = t_start - 1 - s;
    TRY(fill_t_launch(u.t_dev, t, B, st));
    TRY(run_plan(h, pl, st));
    const float* nz = nullptr;
    if (t > 0 && noise) {
      LADIFF_REQUIRE(k < n_noise, LADIFF_ERR_ARG, "ddpm: pre-drawn noise exhausted at step %d (have %lld)", s, (long long)n_noise);
      nz = noise + (size_t)k * B * 128 * L;
      ++k;
    }
    TRY(ddpm_step_launch(u.eps, x, nz, seed, t, h->clip_offset, ddpm_coef(h->un, t), xs, B, 128, L, st));
    h->launches += 2;
  }
  return 0;
}

// ddim_sample (ddpm_loss.py:268-303) over the host array times[0..n_pairs] (decreasing; times[n_pairs] may be -1)
int ddim_run(H* h, Plan* pl, float* x, const float* cond, const int32_t* times, int n_pairs, float eta, const float* noise, int64_t n_noise,
             uint64_t seed, int B, int L, int F, cudaStream_t st) {
  UnetBufs& u = pl->bufs;
  for (int i = 0; i < n_pairs; ++i)
    LADIFF_REQUIRE(times[i] >= 0 && times[i] < kTimesteps && times[i + 1] >= -1 && times[i + 1] < times[i], LADIFF_ERR_ARG,
                   "ddim: times must decrease inside [0, %d) (pair %d: %d -> %d)", kTimesteps, i, times[i], times[i + 1]);
  TRY(prepare_cond(h, pl, cond, B, L, F, st));
  ClView xs = view(u.xin, L, 256, 128, 128);
  TRY(ncl_to_cl_launch(x, nullptr, xs, B, 128, L, st));
  h->launches++;
  int64_t k = 0;
  for (int i = 0; i < n_pairs; ++i) {
    const int t = times[i], tn = times[i + 1];
    TRY(fill_t_launch(u.t_dev, t, B, st));
    TRY(run_plan(h, pl, st));
    const float* nz = nullptr;
    if (tn >= 0 && noise) {       // the reference draws randn_like for every pair with time_next >= 0, even at eta = 0
      LADIFF_REQUIRE(k < n_noise, LADIFF_ERR_ARG, "ddim: pre-drawn noise exhausted at pair %d (have %lld)", i, (long long)n_noise);
      nz = noise + (size_t)k * B * 128 * L;
      ++k;
    }
    TRY(ddpm_step_launch(u.eps, x, nz, seed, t, h->clip_offset, ddim_coef(h->un, t, tn, eta), xs, B, 128, L, st));
    h->launches += 2;
  }
  return 0;
}

}  // namespace

// ================================================================================================ C-ABI
extern "C" int32_t ladiff_create(const LadiffConfig* cfg, LadiffHandle** out) {
  LADIFF_REQUIRE(cfg && out, LADIFF_ERR_ARG, "ladiff_create: null argument");
  LADIFF_REQUIRE(cfg->n_enc_ratios >= 1 && cfg->n_enc_ratios <= LADIFF_MAX_RATIOS && cfg->n_upsampling_ratios >= 0 &&
                     cfg->n_upsampling_ratios <= LADIFF_MAX_RATIOS,
                 LADIFF_ERR_ARG, "ladiff_create: bad ratio counts");
  LADIFF_REQUIRE(cfg->lstm_layers >= 0 && cfg->lstm_layers <= 4, LADIFF_ERR_ARG, "ladiff_create: lstm_layers=%d", cfg->lstm_layers);
  LADIFF_REQUIRE(!cfg->quantization || (cfg->n_q >= 1 && cfg->n_q_used >= 1 && cfg->n_q_used <= cfg->n_q), LADIFF_ERR_ARG,
                 "ladiff_create: n_q=%d n_q_used=%d", cfg->n_q, cfg->n_q_used);
  LadiffHandle* h = new LadiffHandle();
  h->cfg = *cfg;
  h->enc_hop = 1;
  for (int i = 0; i < cfg->n_enc_ratios; ++i) h->enc_hop *= cfg->enc_ratios[i];
  build_keys(h);
  h->dev.assign(h->keys.size(), nullptr);
  h->loaded.assign(h->keys.size(), 0);
  *out = h;
  return 0;
}

extern "C" int32_t ladiff_destroy(LadiffHandle* h) {
  if (!h) return 0;
  for (float* p : h->dev) if (p) cudaFree(p);
  for (void* p : h->owned) cudaFree(p);
  for (Plan* p : h->plans) delete p;
  delete h;
  return 0;
}

extern "C" int32_t ladiff_expected_keys(const LadiffHandle* h) { return h ? (int32_t)h->keys.size() : 0; }
extern "C" int32_t ladiff_expected_key_at(const LadiffHandle* h, int32_t i, const char** name, int64_t* shape4, int32_t* ndim) {
  LADIFF_REQUIRE(h && i >= 0 && i < (int)h->keys.size(), LADIFF_ERR_ARG, "ladiff_expected_key_at: index %d", i);
  const KeySpec& k = h->keys[i];
  if (name) *name = k.name.c_str();
  if (ndim) *ndim = (int32_t)k.shape.size();
  if (shape4) for (size_t d = 0; d < k.shape.size() && d < 4; ++d) shape4[d] = k.shape[d];
  return 0;
}

extern "C" int32_t ladiff_load_weight(LadiffHandle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  LADIFF_REQUIRE(h && name && shape, LADIFF_ERR_ARG, "ladiff_load_weight: null argument");
  LADIFF_REQUIRE(!h->finalized, LADIFF_ERR_STATE, "ladiff_load_weight after ladiff_finalize");
  auto it = h->key_index.find(name);
  LADIFF_REQUIRE(it != h->key_index.end(), LADIFF_ERR_KEY, "unexpected key in state_dict: %s", name);
  const KeySpec& k = h->keys[it->second];
  bool same = (int)k.shape.size() == ndim;
  for (int d = 0; same && d < ndim; ++d) same = k.shape[d] == shape[d];
  LADIFF_REQUIRE(same, LADIFF_ERR_KEY, "size mismatch for %s", name);
  if (!k.alias) {
    LADIFF_REQUIRE(data != nullptr, LADIFF_ERR_ARG, "ladiff_load_weight(%s): null data", name);
    if (!h->dev[it->second]) LADIFF_CUDA_OK(cudaMalloc((void**)&h->dev[it->second], sizeof(float) * (size_t)k.numel()));
    LADIFF_CUDA_OK(cudaMemcpy(h->dev[it->second], data, sizeof(float) * (size_t)k.numel(), cudaMemcpyDefault));
  }
  h->loaded[it->second] = 1;
  return 0;
}

extern "C" int32_t ladiff_finalize(LadiffHandle* h) {
  LADIFF_REQUIRE(h, LADIFF_ERR_ARG, "ladiff_finalize: null handle");
  LADIFF_REQUIRE(!h->finalized, LADIFF_ERR_STATE, "ladiff_finalize called twice");
  for (size_t i = 0; i < h->keys.size(); ++i)
    LADIFF_REQUIRE(h->loaded[i], LADIFF_ERR_KEY, "missing key in state_dict: %s", h->keys[i].name.c_str());
  TRY(fold_codec(h));
  if (h->cfg.run_diff) TRY(fold_unet(h));
  LADIFF_CUDA_OK(cudaDeviceSynchronize());
  h->finalized = true;
  return 0;
}

extern "C" int64_t ladiff_workspace_bytes(const LadiffHandle* h, int32_t B, int32_t T) {
  if (!h || B <= 0 || T <= 0) return 0;
  const int L = T / h->enc_hop;
  size_t stage = codec_ws_bytes(h, B, T);
  // the conditioning codec always runs at hop 320 on T samples; its own handle reports its own need
  if (h->cfg.run_diff) { const size_t u = unet_ws_bytes(h, B, L); if (u > stage) stage = u; }
  return (int64_t)(persist_bytes(B, T, L > 0 ? L : 1) + stage);
}

extern "C" int64_t ladiff_synthesize_workspace_bytes(const LadiffHandle* m, const LadiffHandle* cm, int32_t B, int32_t T) {
  if (!m || !cm || B <= 0 || T <= 0) return 0;
  const int L = T / m->enc_hop;
  size_t stage = codec_ws_bytes(cm, B, T);
  stage = std::max(stage, codec_ws_bytes(m, B, T));
  if (m->cfg.run_diff && L > 0) stage = std::max(stage, unet_ws_bytes(m, B, L));
  return (int64_t)(persist_bytes(B, T, L > 0 ? L : 1) + stage);
}

#define STAGE_PROLOGUE(hh)                                                                      \
  LADIFF_REQUIRE((hh) != nullptr, LADIFF_ERR_ARG, "null handle");                                \
  LADIFF_REQUIRE((hh)->finalized, LADIFF_ERR_STATE, "handle not finalized (call ladiff_finalize)"); \
  cudaStream_t st = (cudaStream_t)stream;

extern "C" int32_t ladiff_encode(LadiffHandle* h, const float* wav, int32_t B, int32_t T, float* z, void* ws, int64_t ws_bytes,
                                 void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(wav && z && B > 0 && T > 0 && T % h->enc_hop == 0, LADIFF_ERR_ARG, "ladiff_encode: T=%d must be a multiple of hop %d", T,
                 h->enc_hop);
  TRY(check_ws(ws, ws_bytes, codec_ws_bytes(h, B, T)));
  return run_encoder(h, wav, B, T, z, Bump(ws), st);
}

extern "C" int32_t ladiff_rvq_encode(LadiffHandle* h, const float* z, int32_t n_q, int32_t B, int32_t F, int64_t* codes, float* quantized,
                                     void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(h->cfg.quantization, LADIFF_ERR_STATE, "this model has no quantizer");
  LADIFF_REQUIRE(z && n_q >= 1 && n_q <= h->cfg.n_q && B > 0 && F > 0, LADIFF_ERR_ARG, "ladiff_rvq_encode: n_q=%d", n_q);
  h->launches++;
  return rvq_encode_launch(z, h->embed, h->embed_t, h->embed_sq, n_q, kBins, h->cfg.rep_dims, B, F, quantized, (long long*)codes, st);
}

extern "C" int32_t ladiff_rvq_decode(LadiffHandle* h, const int64_t* codes, int32_t n_q, int32_t B, int32_t F, float* quantized,
                                     void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(h->cfg.quantization, LADIFF_ERR_STATE, "this model has no quantizer");
  LADIFF_REQUIRE(codes && quantized && n_q >= 1 && n_q <= h->cfg.n_q && B > 0 && F > 0, LADIFF_ERR_ARG, "ladiff_rvq_decode: n_q=%d", n_q);
  h->launches++;
  return rvq_decode_launch((const long long*)codes, h->embed, n_q, kBins, h->cfg.rep_dims, B, F, quantized, st);
}

extern "C" int32_t ladiff_get_cond(LadiffHandle* h, const float* wav, int32_t B, int32_t T, float* cond, int64_t* codes, float* enc_out,
                                   void* ws, int64_t ws_bytes, void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(wav && cond && B > 0 && T > 0 && T % h->enc_hop == 0, LADIFF_ERR_ARG, "ladiff_get_cond: T=%d must be a multiple of hop %d",
                 T, h->enc_hop);
  const int F = T / h->enc_hop;
  const size_t zbytes = align_up(sizeof(float) * (size_t)B * h->cfg.rep_dims * F, 1024);
  TRY(check_ws(ws, ws_bytes, codec_ws_bytes(h, B, T)));
  float* z = enc_out ? enc_out : reinterpret_cast<float*>(ws);
  TRY(run_encoder(h, wav, B, T, z, Bump((char*)ws + zbytes), st));
  if (!h->cfg.quantization) {   // model.py:227 — no quantizer: cond is the encoder output
    if (z != cond) LADIFF_CUDA_OK(cudaMemcpyAsync(cond, z, sizeof(float) * (size_t)B * h->cfg.rep_dims * F, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  h->launches++;
  return rvq_encode_launch(z, h->embed, h->embed_t, h->embed_sq, h->cfg.n_q_used, kBins, h->cfg.rep_dims, B, F, cond, (long long*)codes, st);
}

extern "C" int32_t ladiff_upsample_layer(LadiffHandle* h, int32_t i, const float* x, int32_t B, int32_t Lin, float* y, void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(h->cfg.run_diff && i >= 0 && i < (int)h->un.cond_up.size(), LADIFF_ERR_ARG, "ladiff_upsample_layer: index %d", i);
  LADIFF_REQUIRE(x && y && B > 0 && Lin > 0, LADIFF_ERR_ARG, "ladiff_upsample_layer: bad arguments");
  return run_convtr(h, h->un.cond_up[i], x, Lin, y, 0, false, B, st);
}

extern "C" int32_t ladiff_unet_forward(LadiffHandle* h, const float* x, const int64_t* time, const float* cond, int32_t B, int32_t L,
                                       int32_t F, float* eps, void* ws, int64_t ws_bytes, void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(h->cfg.run_diff, LADIFF_ERR_STATE, "this model has no diffusion UNet (run_diff=False)");
  LADIFF_REQUIRE(x && time && cond && eps && B > 0, LADIFF_ERR_ARG, "ladiff_unet_forward: null argument");
  TRY(check_ws(ws, ws_bytes, unet_ws_bytes(h, B, L)));
  Plan* pl = nullptr;
  TRY(build_plan(h, ws, B, L, st, &pl));
  UnetBufs& u = pl->bufs;
  TRY(prepare_cond(h, pl, cond, B, L, F, st));
  TRY(ncl_to_cl_launch(x, nullptr, view(u.xin, L, 256, 128, 128), B, 128, L, st));
  TRY(time_to_int_launch((const long long*)time, u.t_dev, B, st));
  TRY(run_plan(h, pl, st));
  TRY(cl_to_ncl_f32_launch(u.eps, eps, B, 128, L, st));
  h->launches += 3;
  return 0;
}

extern "C" int32_t ladiff_ddpm_steps(LadiffHandle* h, float* x, const float* cond, const float* noise, int64_t n_noise, uint64_t seed,
                                     int32_t t_start, int32_t n_steps, int32_t B, int32_t L, int32_t F, void* ws, int64_t ws_bytes,
                                     void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(h->cfg.run_diff, LADIFF_ERR_STATE, "this model has no diffusion UNet (run_diff=False)");
  LADIFF_REQUIRE(x && cond && B > 0, LADIFF_ERR_ARG, "ladiff_ddpm_steps: null argument");
  TRY(check_ws(ws, ws_bytes, unet_ws_bytes(h, B, L)));
  Plan* pl = nullptr;
  TRY(build_plan(h, ws, B, L, st, &pl));
  return ddpm_run(h, pl, x, cond, noise, n_noise, seed, t_start, n_steps, B, L, F, st);
}

extern "C" int32_t ladiff_set_clip_offset(LadiffHandle* h, uint64_t clip_offset) {
  LADIFF_REQUIRE(h, LADIFF_ERR_ARG, "null handle");
  h->clip_offset = clip_offset;
  return 0;
}

extern "C" int32_t ladiff_ddim_steps(LadiffHandle* h, float* x, const float* cond, const int32_t* times, int32_t n_pairs, double eta,
                                     const float* noise, int64_t n_noise, uint64_t seed, int32_t B, int32_t L, int32_t F, void* ws,
                                     int64_t ws_bytes, void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(h->cfg.run_diff, LADIFF_ERR_STATE, "this model has no diffusion UNet (run_diff=False)");
  LADIFF_REQUIRE(x && cond && times && n_pairs >= 0 && B > 0, LADIFF_ERR_ARG, "ladiff_ddim_steps: bad argument");
  TRY(check_ws(ws, ws_bytes, unet_ws_bytes(h, B, L)));
  Plan* pl = nullptr;
  TRY(build_plan(h, ws, B, L, st, &pl));
  return ddim_run(h, pl, x, cond, times, n_pairs, (float)eta, noise, n_noise, seed, B, L, F, st);
}

extern "C" int32_t ladiff_randn(LadiffHandle* h, float* x, int32_t B, int64_t n_per_clip, uint64_t seed, int32_t uniform, void* stream) {
  LADIFF_REQUIRE(h && x && B > 0 && n_per_clip > 0, LADIFF_ERR_ARG, "ladiff_randn: bad argument");
  h->launches++;
  return randn_fill_launch(x, (long long)B * n_per_clip, seed, h->clip_offset * (unsigned long long)n_per_clip, uniform ? 1 : 0, (cudaStream_t)stream);
}

extern "C" int32_t ladiff_q_sample(LadiffHandle* h, const float* x_start, const int64_t* t, const float* noise, float* out, int32_t B,
                                   int64_t n_per_clip, void* ws, int64_t ws_bytes, void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(h->cfg.run_diff, LADIFF_ERR_STATE, "this model has no diffusion process (run_diff=False)");
  LADIFF_REQUIRE(x_start && t && noise && out && B > 0 && n_per_clip > 0, LADIFF_ERR_ARG, "ladiff_q_sample: bad argument");
  TRY(check_ws(ws, ws_bytes, 1024 + sizeof(int) * (size_t)B));
  int* t_dev = reinterpret_cast<int*>(ws);
  TRY(time_to_int_launch((const long long*)t, t_dev, B, st));
  h->launches += 2;
  return q_sample_launch(x_start, noise, t_dev, h->un.d_sqrt_ac, h->un.d_sqrt_1m_ac, out, B, n_per_clip, st);
}

extern "C" int32_t ladiff_axpby(float* x, double a, const float* y, double b, int64_t n, void* stream) {
  LADIFF_REQUIRE(x && n > 0, LADIFF_ERR_ARG, "ladiff_axpby: bad argument");
  return axpby_launch(x, (float)a, y, (float)b, n, (cudaStream_t)stream);
}

extern "C" int32_t ladiff_p_losses(LadiffHandle* h, const float* x_start, const int64_t* t, const float* cond, const float* noise, int32_t B,
                                   int32_t L, int32_t F, float* loss, float* pred_x_start, float* x_t, float* model_out, void* ws,
                                   int64_t ws_bytes, void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(h->cfg.run_diff, LADIFF_ERR_STATE, "this model has no diffusion UNet (run_diff=False)");
  LADIFF_REQUIRE(x_start && t && cond && noise && loss && x_t && B > 0, LADIFF_ERR_ARG, "ladiff_p_losses: null argument");
  TRY(check_ws(ws, ws_bytes, unet_ws_bytes(h, B, L)));
  Plan* pl = nullptr;
  TRY(build_plan(h, ws, B, L, st, &pl));
  UnetBufs& u = pl->bufs;
  TRY(time_to_int_launch((const long long*)t, u.t_dev, B, st));
  TRY(q_sample_launch(x_start, noise, u.t_dev, h->un.d_sqrt_ac, h->un.d_sqrt_1m_ac, x_t, B, (long long)128 * L, st));   // ddpm_loss.py:410
  TRY(prepare_cond(h, pl, cond, B, L, F, st));
  TRY(ncl_to_cl_launch(x_t, nullptr, view(u.xin, L, 256, 128, 128), B, 128, L, st));
  TRY(run_plan(h, pl, st));                                     // model(x, t, cond): :418 and :423 are the same evaluation forward-only
  TRY(p_losses_launch(u.eps, noise, x_t, u.t_dev, h->un.tb.sqrt_recip_ac, h->un.tb.sqrt_recipm1_ac, h->un.d_p2w, pred_x_start, model_out,
                      u.inv_scale, loss, B, 128, L, st));
  h->launches += 5;
  return 0;
}

extern "C" int32_t ladiff_sdsdr(const float* est, const float* target, float* out, int32_t B, int64_t n, double clip_value, void* stream) {
  LADIFF_REQUIRE(est && target && out && B > 0 && n > 1, LADIFF_ERR_ARG, "ladiff_sdsdr: bad argument");
  return sdsdr_launch(est, target, out, B, n, (float)clip_value, (cudaStream_t)stream);
}

extern "C" int32_t ladiff_decode(LadiffHandle* h, const float* z, int32_t B, int32_t L, float* wav, void* ws, int64_t ws_bytes,
                                 void* stream) {
  STAGE_PROLOGUE(h);
  LADIFF_REQUIRE(z && wav && B > 0 && L > 0, LADIFF_ERR_ARG, "ladiff_decode: bad arguments");
  TRY(check_ws(ws, ws_bytes, codec_ws_bytes(h, B, L * h->enc_hop)));
  return run_decoder(h, z, B, L, wav, Bump(ws), st);
}

extern "C" int32_t ladiff_normalize_clips(float* x, int32_t B, int64_t n, int32_t mode, void* stream) {
  LADIFF_REQUIRE(x && B > 0 && n > 1 && mode >= 0 && mode <= 2, LADIFF_ERR_ARG, "ladiff_normalize_clips: bad arguments");
  return normalize_clips_launch(x, B, n, mode, (cudaStream_t)stream);
}

// common tail of the synthesis entry points: cond [B][128][F] (persist region) -> wav_out
struct SamplerSpec {
  int kind = 0;                 // 0: halfway_sampling(t = n_steps) from the upsampled cond; 1: ddim_sample from noise; 2: p_sample_loop from noise
  int n_steps = 0;
  std::vector<int32_t> times;   // DDIM time pairs
  float eta = 0.f;
  const float* init = nullptr;  // initial noise for kinds 1, 2 (null: in-kernel generator)
};
static int synth_from_cond(LadiffHandle* m, const float* cond, int B, int T, const SamplerSpec& sp, const float* noise, int64_t n_noise,
                           uint64_t seed, float* wav_out, float* latent_out, Bump bp, void* scratch, cudaStream_t st) {
  const int L = T / m->enc_hop, F = T / 320;
  float* x = latent_out ? latent_out : bp.get<float>((size_t)B * 128 * L);
  float* ta = bp.get<float>((size_t)B * 128 * L);
  float* tb = bp.get<float>((size_t)B * 128 * L);
  Plan* pl = nullptr;
  if (sp.kind == 0) {
    TRY(run_cond_upsample(m, cond, B, F, x, ta, tb, st));                                               // sample.py:125-128
    TRY(normalize_clips_launch(x, B, (long long)128 * L, 0, st));                                       // :129
    TRY(build_plan(m, scratch, B, L, st, &pl));
    TRY(ddpm_run(m, pl, x, cond, noise, n_noise, seed, sp.n_steps, sp.n_steps, B, L, F, st));           // :130
  } else {
    if (sp.init) LADIFF_CUDA_OK(cudaMemcpyAsync(x, sp.init, sizeof(float) * (size_t)B * 128 * L, cudaMemcpyDeviceToDevice, st));
    else TRY(randn_fill_launch(x, (long long)B * 128 * L, seed, m->clip_offset * (unsigned long long)128 * L, 0, st));   // torch.randn(shape), ddpm_loss.py:256,277
    TRY(build_plan(m, scratch, B, L, st, &pl));
    if (sp.kind == 1) TRY(ddim_run(m, pl, x, cond, sp.times.data(), (int)sp.times.size() - 1, sp.eta, noise, n_noise, seed, B, L, F, st));
    else TRY(ddpm_run(m, pl, x, cond, noise, n_noise, seed, kTimesteps, sp.n_steps, B, L, F, st));
  }
  TRY(run_decoder(m, x, B, L, wav_out, Bump(scratch), st));                                             // :131
  TRY(normalize_clips_launch(wav_out, B, T, 1, st));                                                    // :133-134
  m->launches += 2;
  return 0;
}

static int synth_common(LadiffHandle* m, LadiffHandle* cm, const float* wav_in, int32_t B, int32_t T, const SamplerSpec& sp, const float* noise,
                        int64_t n_noise, uint64_t seed, float* wav_out, float* latent_out, void* ws, int64_t ws_bytes, void* stream) {
  STAGE_PROLOGUE(m);
  LADIFF_REQUIRE(cm && cm->finalized, LADIFF_ERR_STATE, "conditioning model not finalized");
  LADIFF_REQUIRE(m->cfg.run_diff && cm->cfg.quantization, LADIFF_ERR_STATE, "synthesize needs a diffusion model and a quantising cond codec");
  LADIFF_REQUIRE(wav_in && wav_out && B > 0 && T > 0 && T % 640 == 0, LADIFF_ERR_ARG,
                 "ladiff_synthesize: T=%d must be a multiple of 640 (sample.py:87)", T);
  LADIFF_REQUIRE(T % m->enc_hop == 0 && T % cm->enc_hop == 0 && cm->enc_hop == 320, LADIFF_ERR_ARG, "T=%d is not a multiple of the codec hops", T);
  const int L = T / m->enc_hop, F = T / cm->enc_hop;
  const size_t pb = persist_bytes(B, T, L);
  TRY(check_ws(ws, ws_bytes, (size_t)ladiff_synthesize_workspace_bytes(m, cm, B, T)));
  Bump bp(ws);
  float* cond = bp.get<float>((size_t)B * 128 * F);
  void* scratch = (char*)ws + pb;
  TRY(ladiff_get_cond(cm, wav_in, B, T, cond, nullptr, nullptr, scratch, ws_bytes - (int64_t)pb, stream));   // sample.py:94
  return synth_from_cond(m, cond, B, T, sp, noise, n_noise, seed, wav_out, latent_out, bp, scratch, st);
}

extern "C" int32_t ladiff_synthesize(LadiffHandle* m, LadiffHandle* cm, const float* wav_in, int32_t B, int32_t T, int32_t n_steps,
                                     const float* noise, int64_t n_noise, uint64_t seed, float* wav_out, float* latent_out, void* ws,
                                     int64_t ws_bytes, void* stream) {
  SamplerSpec sp;
  sp.kind = 0; sp.n_steps = n_steps;
  return synth_common(m, cm, wav_in, B, T, sp, noise, n_noise, seed, wav_out, latent_out, ws, ws_bytes, stream);
}

// ddim_sample's time grid (ddpm_loss.py:273-275): torch.linspace(-1, T-1, steps = S+1) in fp32, truncated to int, reversed.
// torch.linspace computes start + step*i for the first half and end - step*(steps-1-i) for the second half (fp32).
static void ddim_times(int total, int S, std::vector<int32_t>* out) {
  const int steps = S + 1;
  const float start = -1.f, end = (float)(total - 1);
  const float step = (end - start) / (float)(steps - 1);
  std::vector<int32_t> t(steps);
  const int half = steps / 2;
  for (int i = 0; i < steps; ++i) {
    const float v = i < half ? start + step * (float)i : end - step * (float)(steps - 1 - i);
    t[i] = (int32_t)v;      // .int(): truncation toward zero
  }
  out->assign(t.rbegin(), t.rend());
}

// The same per-file body with the reference's DDIM sampler in place of halfway_sampling:
// get_cond → ddim_sample((B, 128, L), condition) with `sampling_timesteps` steps from N(0, I) → decoder → normalise.
// init_noise [B,128,L] optional (null: in-kernel generator); noise [n,B,128,L] one per pair with time_next >= 0 (only used if eta != 0).
extern "C" int32_t ladiff_synthesize_ddim(LadiffHandle* m, LadiffHandle* cm, const float* wav_in, int32_t B, int32_t T,
                                          int32_t sampling_timesteps, double eta, const float* init_noise, const float* noise, int64_t n_noise,
                                          uint64_t seed, float* wav_out, float* latent_out, void* ws, int64_t ws_bytes, void* stream) {
  LADIFF_REQUIRE(sampling_timesteps >= 1 && sampling_timesteps <= kTimesteps, LADIFF_ERR_ARG, "ladiff_synthesize_ddim: sampling_timesteps=%d",
                 sampling_timesteps);
  SamplerSpec sp;
  sp.kind = 1; sp.eta = (float)eta; sp.init = init_noise;
  ddim_times(kTimesteps, sampling_timesteps, &sp.times);
  return synth_common(m, cm, wav_in, B, T, sp, noise, n_noise, seed, wav_out, latent_out, ws, ws_bytes, stream);
}

extern "C" int32_t ladiff_ddim_times(int32_t total_timesteps, int32_t sampling_timesteps, int32_t* times_out) {
  LADIFF_REQUIRE(times_out && sampling_timesteps >= 1 && total_timesteps >= 1, LADIFF_ERR_ARG, "ladiff_ddim_times: bad argument");
  std::vector<int32_t> t;
  ddim_times(total_timesteps, sampling_timesteps, &t);
  for (size_t i = 0; i < t.size(); ++i) times_out[i] = t[i];
  return 0;
}

// Receiver side of the codec: the conditioning arrives as RVQ indices instead of a waveform.
extern "C" int32_t ladiff_synthesize_codes(LadiffHandle* m, LadiffHandle* cm, const int64_t* codes, int32_t n_q, int32_t B, int32_t F,
                                           int32_t n_steps, const float* noise, int64_t n_noise, uint64_t seed, float* wav_out,
                                           float* latent_out, void* ws, int64_t ws_bytes, void* stream) {
  STAGE_PROLOGUE(m);
  LADIFF_REQUIRE(cm && cm->finalized, LADIFF_ERR_STATE, "conditioning model not finalized");
  LADIFF_REQUIRE(m->cfg.run_diff && cm->cfg.quantization, LADIFF_ERR_STATE, "synthesize needs a diffusion model and a quantising cond codec");
  LADIFF_REQUIRE(codes && wav_out && B > 0 && F > 0 && F % 2 == 0 && n_q >= 1 && n_q <= cm->cfg.n_q, LADIFF_ERR_ARG,
                 "ladiff_synthesize_codes: n_q=%d F=%d (F must be even: the script decodes multiples of 640 samples)", n_q, F);
  const int T = F * 320;
  LADIFF_REQUIRE(T % m->enc_hop == 0, LADIFF_ERR_ARG, "T=%d is not a multiple of the decoder hop", T);
  const int L = T / m->enc_hop;
  const size_t pb = persist_bytes(B, T, L);
  TRY(check_ws(ws, ws_bytes, (size_t)ladiff_synthesize_workspace_bytes(m, cm, B, T)));
  Bump bp(ws);
  float* cond = bp.get<float>((size_t)B * 128 * F);
  void* scratch = (char*)ws + pb;
  TRY(ladiff_rvq_decode(cm, codes, n_q, B, F, cond, stream));                                           // vq.py:108-113
  SamplerSpec sp;
  sp.kind = 0; sp.n_steps = n_steps;
  return synth_from_cond(m, cond, B, T, sp, noise, n_noise, seed, wav_out, latent_out, bp, scratch, st);
}

extern "C" int32_t ladiff_set_conv_impl(LadiffHandle* h, int32_t impl) {
  LADIFF_REQUIRE(h && impl >= 0 && impl <= 2, LADIFF_ERR_ARG, "ladiff_set_conv_impl: impl=%d", impl);
  h->conv_impl = impl;
  return 0;
}
extern "C" int32_t ladiff_set_skip_ops(LadiffHandle* h, int32_t mask) {
  LADIFF_REQUIRE(h && mask >= 0 && mask < 32, LADIFF_ERR_ARG, "ladiff_set_skip_ops: mask=%d", mask);
  if (mask != h->skip_ops) {
    LADIFF_CUDA_OK(cudaDeviceSynchronize());
    for (Plan* p : h->plans)
      if (p->gexec) { cudaGraphExecDestroy(p->gexec); p->gexec = nullptr; }       // re-captured with the new op set on the next run
  }
  h->skip_ops = mask;
  return 0;
}
extern "C" int32_t ladiff_set_profiling(LadiffHandle* h, int32_t on) {
  LADIFF_REQUIRE(h, LADIFF_ERR_ARG, "null handle");
  h->profiling = on < 0 ? 0 : (on > 2 ? 2 : on);   // 1: events around conv launches; 2: around every op of the evaluation
  if (!on) h->last_plan = nullptr;
  return 0;
}
// out[0] = summed duration (ms) of the conv launches of the most recent UNet evaluation, out[1] = their algorithmic FLOPs,
// out[2] = number of conv launches, out[3] = duration (ms) of that whole evaluation.  Synchronises the device.
extern "C" int32_t ladiff_profile_report(LadiffHandle* h, double* out4) {
  LADIFF_REQUIRE(h && out4, LADIFF_ERR_ARG, "null argument");
  LADIFF_REQUIRE(h->last_plan != nullptr, LADIFF_ERR_STATE, "no profiled UNet evaluation yet");
  LADIFF_CUDA_OK(cudaDeviceSynchronize());
  Plan* pl = h->last_plan;
  double ms = 0.0, fl = 0.0, n = 0.0;
  for (size_t i = 0; i < pl->ops.size(); ++i) {
    if (pl->op_flops[i] <= 0.0) continue;
    float t = 0.f;
    LADIFF_CUDA_OK(cudaEventElapsedTime(&t, pl->ev[2 * i], pl->ev[2 * i + 1]));
    ms += t; fl += pl->op_flops[i]; n += 1.0;
  }
  float tot = 0.f;
  LADIFF_CUDA_OK(cudaEventElapsedTime(&tot, pl->ev[2 * pl->ops.size()], pl->ev[2 * pl->ops.size() + 1]));
  out4[0] = ms; out4[1] = fl; out4[2] = n; out4[3] = tot;
  return 0;
}
// Per-conv-launch table of the most recent profiled evaluation, one text line per launch: "<ms> <gflop> <label>".
extern "C" int32_t ladiff_profile_dump(LadiffHandle* h, char* buf, int64_t cap) {
  LADIFF_REQUIRE(h && buf && cap > 0, LADIFF_ERR_ARG, "null argument");
  LADIFF_REQUIRE(h->last_plan != nullptr, LADIFF_ERR_STATE, "no profiled UNet evaluation yet");
  LADIFF_CUDA_OK(cudaDeviceSynchronize());
  Plan* pl = h->last_plan;
  std::string out;
  pl->op_label.resize(pl->ops.size());
  for (size_t i = 0; i < pl->ops.size(); ++i) {
    if (pl->op_flops[i] <= 0.0 && h->profiling != 2) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, pl->ev[2 * i], pl->ev[2 * i + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
    char line[256];
    snprintf(line, sizeof(line), "%.4f %.3f %s\n", t, pl->op_flops[i] * 1e-9, pl->op_label[i].c_str());
    out += line;
  }
  snprintf(buf, (size_t)cap, "%s", out.c_str());
  return 0;
}
extern "C" int64_t ladiff_take_launch_count(LadiffHandle* h) {
  if (!h) return 0;
  const long long n = h->launches;
  h->launches = 0;
  return n;
}

extern "C" int32_t ladiff_op_conv1d_cl(const void* x_h16, const float* w, const float* bias, int32_t B, int32_t L, int32_t Cin,
                                       int32_t Cout, int32_t k, void* y, int32_t y_f32, int32_t impl, float* gn_stats) {
  LADIFF_REQUIRE(x_h16 && w && y && Cin % 64 == 0 && Cout % 128 == 0 && k >= 1 && k <= TC_MAX_TAPS && (k & 1), LADIFF_ERR_ARG,
                 "ladiff_op_conv1d_cl: Cin %% 64, Cout %% 128, odd k <= %d required", TC_MAX_TAPS);
  LADIFF_REQUIRE(impl >= 0 && impl <= 6, LADIFF_ERR_ARG, "ladiff_op_conv1d_cl: impl=%d", impl);
  h16* wp = nullptr; float2* stats = nullptr;
  LADIFF_CUDA_OK(cudaMalloc((void**)&wp, sizeof(h16) * (size_t)Cout * Cin * k));
  int rc = pack_conv_launch(w, wp, Cout, Cin, k, 0, 0);
  CUtensorMap tmW;
  if (!rc) rc = tc_make_tmap_w(&tmW, wp, Cout, Cin * k);
  TcConvDesc d;
  memset(&d, 0, sizeof(d));
  d.kind = TC_KIND_PLAIN; d.Cin = Cin; d.K = k; d.CoutV = Cout; d.w = wp; d.tmW = &tmW; d.Ktot = Cin * k; d.bias = bias;
  d.x = (const h16*)x_h16; d.x_bstride = (long long)L * Cin; d.x_pitch = Cin; d.Lin = L;
  if (y_f32) d.out32 = (float*)y;
  else { d.out = (h16*)y; d.out_bstride = (long long)L * Cout; d.out_pitch = Cout; }
  d.B = B; d.tap_share = impl == 2 ? 0 : 1; d.want_transposed = impl == 3 ? 1 : (impl == 4 ? 2 : 0);
  if (impl == 5) { d.want_two_per_sm = 1; d.want_nt = 128; }
  if (impl == 6) d.want_pair = 1;           // CTA pairs along N, weight tiles by TMA multicast
  TcConvParams p;
  TcRefView rv;
  if (!rc) rc = tc_conv_plan(d, &p, &rv);
  if (!rc && gn_stats) {
    if (cudaMalloc((void**)&stats, sizeof(float2) * (size_t)B * p.n_ptiles * p.stat_parts * (Cout / 32)) != cudaSuccess) rc = LADIFF_ERR_CUDA;
    else cudaMemset(stats, 0, sizeof(float2) * (size_t)B * p.n_ptiles * p.stat_parts * (Cout / 32));
    p.stats = stats;
  }
  if (!rc) rc = impl == 1 ? tc_conv_ref_launch(p, rv, 0) : tc_conv_launch(p, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (!rc && e != cudaSuccess) { ladiff_set_error("ladiff_op_conv1d_cl: %s", cudaGetErrorString(e)); rc = LADIFF_ERR_CUDA; }
  if (!rc && gn_stats) {   // reduce the per-tile partials on the host: [B][Cout/32][2]
    const int npt = p.n_ptiles * p.stat_parts;
    std::vector<float2> hst((size_t)B * npt * (Cout / 32));
    cudaMemcpy(hst.data(), stats, sizeof(float2) * hst.size(), cudaMemcpyDeviceToHost);
    std::vector<float> red((size_t)B * (Cout / 32) * 2, 0.f);
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < npt; ++t)
        for (int s = 0; s < Cout / 32; ++s) {
          const float2 v = hst[((size_t)b * npt + t) * (Cout / 32) + s];
          red[((size_t)b * (Cout / 32) + s) * 2] += v.x;
          red[((size_t)b * (Cout / 32) + s) * 2 + 1] += v.y;
        }
    cudaMemcpy(gn_stats, red.data(), sizeof(float) * red.size(), cudaMemcpyDefault);
  }
  cudaFree(wp);
  if (stats) cudaFree(stats);
  return rc;
}

// Operator-level timing for kernel tuning (profiles/conv_sweep.py): plans once, launches `iters` times back to back on the legacy
// stream between two CUDA events (after `warm` untimed launches); ms_out[0] = mean duration of one launch.  kind: TC_KIND_*;
// want_nt / want_nclip / want_two / want_t: tile-shape overrides (0 = cost model); stats != 0: GroupNorm partials on.
extern "C" int32_t ladiff_op_conv1d_bench(const void* x_h16, const float* w, const float* bias, int32_t B, int32_t L, int32_t Cin, int32_t Cout,
                                          int32_t k, int32_t want_nt, int32_t want_nclip, int32_t want_two, int32_t want_t, int32_t stats_on,
                                          int32_t warm, int32_t iters, float* ms_out, char* label, int32_t label_cap) {
  LADIFF_REQUIRE(x_h16 && w && ms_out && Cin % 64 == 0 && Cout % 128 == 0 && k >= 1 && k <= TC_MAX_TAPS && (k & 1), LADIFF_ERR_ARG,
                 "ladiff_op_conv1d_bench: bad argument");
  h16 *wp = nullptr, *y = nullptr; float2* stats = nullptr;
  LADIFF_CUDA_OK(cudaMalloc((void**)&wp, sizeof(h16) * (size_t)Cout * Cin * k));
  LADIFF_CUDA_OK(cudaMalloc((void**)&y, sizeof(h16) * (size_t)B * L * Cout));
  int rc = pack_conv_launch(w, wp, Cout, Cin, k, 0, 0);
  CUtensorMap tmW;
  if (!rc) rc = tc_make_tmap_w(&tmW, wp, Cout, Cin * k);
  TcConvDesc d;
  memset(&d, 0, sizeof(d));
  d.kind = TC_KIND_PLAIN; d.Cin = Cin; d.K = k; d.CoutV = Cout; d.w = wp; d.tmW = &tmW; d.Ktot = Cin * k; d.bias = bias;
  d.x = (const h16*)x_h16; d.x_bstride = (long long)L * Cin; d.x_pitch = Cin; d.Lin = L;
  d.out = y; d.out_bstride = (long long)L * Cout; d.out_pitch = Cout;
  d.B = B; d.tap_share = 1; d.want_nt = want_nt; d.want_nclip = want_nclip; d.want_two_per_sm = want_two == 1 ? 1 : 0; d.want_transposed = want_t;
  d.want_pair = want_two == 2 ? 1 : 0;      // want_two: 1 = two CTAs per SM, 2 = multicast CTA pairs
  TcConvParams p;
  if (!rc) rc = tc_conv_plan(d, &p, nullptr);
  if (!rc && stats_on) {
    const size_t n = (size_t)B * p.n_ptiles * p.stat_parts * (Cout / 32);
    if (cudaMalloc((void**)&stats, sizeof(float2) * n) != cudaSuccess) rc = LADIFF_ERR_CUDA;
    p.stats = stats;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (!rc && (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)) rc = LADIFF_ERR_CUDA;
  for (int i = 0; !rc && i < warm; ++i) rc = tc_conv_launch(p, 0);
  if (!rc) cudaEventRecord(e0, 0);
  for (int i = 0; !rc && i < iters; ++i) rc = tc_conv_launch(p, 0);
  if (!rc) {
    cudaEventRecord(e1, 0);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess) { ladiff_set_error("ladiff_op_conv1d_bench: %s", cudaGetErrorString(e)); rc = LADIFF_ERR_CUDA; }
    ms_out[0] = ms / (float)(iters > 0 ? iters : 1);
  }
  if (label && label_cap > 0)
    snprintf(label, (size_t)label_cap, "NT=%d nclip=%d tiles=%d S=%d a_cap=%d CR=%d minb=%d posM=%d NCH=%d pair=%d", p.NT, p.NCLIP,
             p.transposed ? p.n_chtiles * p.n_ntiles : p.MT * p.n_ntiles, p.S, p.a_cap, p.CR, p.minb, p.transposed, p.NCH, p.pair);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaDeviceSynchronize();
  cudaFree(wp); cudaFree(y);
  if (stats) cudaFree(stats);
  return rc;
}

// Operator-level entry for the mid-block attention core (tests): qkv [B][L][384] 16-bit (q | k | v, 4 heads x 32) -> out [B][L][128].
// impl 0 = what the UNet would pick, 1 = tiled SIMT, 2 = SIMT with keys in shared memory, 3 = tcgen05.  Synchronises.
extern "C" int32_t ladiff_op_fullattn(const void* qkv_h16, void* out_h16, int32_t B, int32_t L, int32_t impl) {
  LADIFF_REQUIRE(qkv_h16 && out_h16 && B > 0 && L > 0 && impl >= 0 && impl <= 3, LADIFF_ERR_ARG, "ladiff_op_fullattn: bad argument");
  ClView q = view((h16*)qkv_h16, L, 384, 384), o = view((h16*)out_h16, L, 128, 128);
  int rc = impl == 0 ? fullattn_launch(q, o, B, L, 0) : fullattn_launch_impl(q, o, B, L, impl, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (!rc && e != cudaSuccess) { ladiff_set_error("ladiff_op_fullattn: %s", cudaGetErrorString(e)); rc = LADIFF_ERR_CUDA; }
  return rc;
}
