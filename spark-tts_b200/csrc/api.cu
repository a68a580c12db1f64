// C ABI of libsparkcodec: handle, checkpoint ingestion, workspace planning and the detokenize schedule.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost a predicted branch when no tool is attached

#include "common.cuh"
#include "gemm_params.cuh"
#include "pack.h"

namespace sparkcodec {

static thread_local char g_err[1024] = "";
thread_local int64_t* g_launch_counter = nullptr;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Every C-ABI entry point that touches the handle's device switches to it for the duration of the call only: the
// caller's current device is restored on every exit path (a PyTorch process that drives several GPUs keeps its
// own notion of "current device").
struct DeviceGuard {
  int prev = -1, target;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) : target(dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != target) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define SC_ON_DEVICE(dev)                 \
  ::sparkcodec::DeviceGuard _dev_guard(dev); \
  SC_CUDA(_dev_guard.err)

// NVTX range per stage of a pass (token stages, each backbone, conv-in, each up-block, head): what a timeline tool
// (Nsight Systems, ncu --nvtx) groups the launches by.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  explicit NvtxRange(const std::string& name) { nvtxRangePushA(name.c_str()); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const { int64_t n = 1; for (auto d : shape) n *= d; return n; }
};

struct SnakeParams { float* alpha = nullptr; float* inv = nullptr; };

struct ConvNeXt {
  float* dw_w = nullptr;   // [7][C]
  float* dw_b = nullptr;
  float* ln_w = nullptr;   // plain LayerNorm affine (null when AdaLN)
  float* ln_b = nullptr;
  GemmWeights pw1, pw2;    // pw2 has gamma folded in
};
struct Backbone {
  bool ada = false;
  GemmWeights embed;
  float* norm_w = nullptr; float* norm_b = nullptr;
  std::vector<ConvNeXt> blocks;
  float* final_w = nullptr; float* final_b = nullptr;   // x3 of the following SamplingBlock folded in
};
struct ResUnit {
  GemmWeights c7, c1;
  SnakeParams s_mid;    // snake between conv7 and conv1 (applied in conv7's epilogue)
  SnakeParams s_next;   // snake that consumes this unit's output (applied in conv1's epilogue)
};
struct UpBlock {
  int stride = 1, c_in = 0, c_out = 0;
  GemmWeights convt;
  SnakeParams s_after;  // first residual unit's input snake, phase-replicated (convT epilogue)
  ResUnit ru[3];
};

// ---- speaker half of tokenize: ECAPA-TDNN trunk, perceiver resampler, FSQ project_in (all fp32 on the device) ----
struct SpkConv {                 // Conv1d (+ ReLU + BatchNorm(eval) folded into scale / shift)
  float *w = nullptr, *b = nullptr, *scale = nullptr, *shift = nullptr;
  int c_out = 0, c_in = 0, k = 1;
};
struct SpkRes2Block {
  SpkConv c0, c2;
  std::vector<SpkConv> convs;    // Res2Conv1dReluBn: scale - 1 convs of `width` channels
  float *se_w1 = nullptr, *se_b1 = nullptr, *se_w2 = nullptr, *se_b2 = nullptr;
  int dilation = 1, se_dim = 0;
};
struct SpkPercLayer {
  float *wq = nullptr, *wkv = nullptr, *wo = nullptr, *w_ff0 = nullptr, *b_ff0 = nullptr, *w_ff2 = nullptr, *b_ff2 = nullptr;
};
struct SpeakerModel {
  bool present = false;
  // mel front end (defaults: the released config.yaml mel_params; sparkcodec_set_mel_params overrides before finalize)
  int sample_rate = 16000, n_fft = 1024, win = 640, hop = 320, n_mels = 128;
  float fmin = 10.f, fmax = -1.f;            // fmax < 0: sample_rate / 2
  float *dft = nullptr, *fb = nullptr;       // (2 * bins, win) windowed DFT rows [re | im]; (n_mels, bins) filterbank
  int bins = 0;
  SpkConv layer1, conv_cat;
  std::vector<SpkRes2Block> blocks;
  int channels = 0, width = 0;
  float *latents = nullptr, *pc_w = nullptr, *pc_b = nullptr, *norm_gamma = nullptr, *fsq_win = nullptr, *fsq_bin = nullptr;
  std::vector<SpkPercLayer> layers;
  int dim = 0, n_lat = 0, heads = 8, dim_head = 64, ff_inner = 0;
};

}  // namespace sparkcodec

using namespace sparkcodec;

struct sparkcodec_handle {
  sparkcodec_config cfg;
  int device = 0, num_sms = 148;
  bool finalized = false;
  int impl = SPARKCODEC_IMPL_TC;
  int64_t launches = 0;
  std::map<std::string, HostTensor> host;
  std::vector<void*> allocs;
  int* err_flag = nullptr;   // device int[4]
  bool profile = false;      // record a CUDA-event pair around every launch (bench.py roofline pass)
  struct ProfRec { std::string name; double flops, bytes; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof;

  // token stages
  float *codebook = nullptr, *vq_mat = nullptr, *vq_vec = nullptr, *vq_w = nullptr, *vq_b = nullptr;
  int* fsq_levels = nullptr;
  float *fsq_wpo = nullptr, *fsq_bpo = nullptr, *spk_w = nullptr, *spk_b = nullptr;
  float *ada_w = nullptr, *ada_b = nullptr;
  int n_ada = 0;
  // prenet
  std::vector<Backbone> backbones;
  GemmWeights linear;
  // encode side (optional: only when the checkpoint carries encoder.* and quantizer.in_project.*)
  bool has_encoder = false;
  std::vector<Backbone> enc_backbones;
  float *tok_mat = nullptr, *tok_vec = nullptr;        // in_project . encoder.project folded: (codebook_dim, C)
  float *codes_n = nullptr, *codes_sq = nullptr;       // F.normalize(codebook) rows and their squared norms
  SpeakerModel spk;                                    // optional: only when the checkpoint carries the ECAPA / perceiver tensors
  // wave generator
  GemmWeights conv_in;
  SnakeParams s_conv_in;   // first block's input snake (conv_in epilogue)
  std::vector<UpBlock> ups;
  SnakeParams s_head;
  float* head_w = nullptr;   // [7][C]
  float head_bias = 0.f;
  int head_c = 0;
  int hop = 1;
};

namespace sparkcodec {

// ------------------------------------------------------------------------------- device helpers
static int dev_alloc(sparkcodec_handle* h, size_t bytes, void** out) {
  void* p = nullptr;
  SC_CUDA(cudaMalloc(&p, bytes ? bytes : 4));
  h->allocs.push_back(p);
  *out = p;
  return 0;
}
template <typename T>
static int upload(sparkcodec_handle* h, const std::vector<T>& v, T** out) {
  void* p;
  SC_TRY(dev_alloc(h, v.size() * sizeof(T), &p));
  SC_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<T*>(p);
  return 0;
}
static int upload_gemm(sparkcodec_handle* h, const PackedGemm& pk, GemmWeights* g) {
  g->c_in = pk.c_in; g->n_total = pk.n_total; g->kt = pk.kt; g->taps = pk.taps;
  uint16_t *hi, *lo;
  SC_TRY(upload(h, pk.w_hi, &hi));
  SC_TRY(upload(h, pk.w_lo, &lo));
  g->w_hi = reinterpret_cast<__nv_bfloat16*>(hi);
  g->w_lo = reinterpret_cast<__nv_bfloat16*>(lo);
  if (pk.w_h16.empty()) { set_error("dense layer with K = %d x %d: not a multiple of 32", pk.kt, pk.c_in); return SPARKCODEC_EINVAL; }
  SC_TRY(upload(h, pk.w_h16, &hi));
  SC_TRY(upload(h, pk.w_p8, &lo));
  g->w_h16 = reinterpret_cast<__nv_bfloat16*>(hi);
  g->w_p8 = reinterpret_cast<__nv_bfloat16*>(lo);
  SC_TRY(upload(h, pk.bias, &g->bias));
  return make_weight_tmaps(*g);
}

static int get(sparkcodec_handle* h, const std::string& key, const HostTensor** out,
               std::initializer_list<int64_t> shape = {}) {
  auto it = h->host.find(key);
  if (it == h->host.end()) {
    set_error("missing checkpoint tensor '%s'", key.c_str());
    return SPARKCODEC_EMISSING;
  }
  if (shape.size()) {
    std::vector<int64_t> want(shape);
    if (it->second.shape != want) {
      std::string got;
      for (auto d : it->second.shape) got += std::to_string(d) + ",";
      std::string exp;
      for (auto d : want) exp += std::to_string(d) + ",";
      set_error("tensor '%s' has shape (%s) but the config implies (%s)", key.c_str(), got.c_str(), exp.c_str());
      return SPARKCODEC_EINVAL;
    }
  }
  *out = &it->second;
  return 0;
}

// weight of a (possibly weight-normed) conv: `prefix.weight` or fold(prefix.weight_g, prefix.weight_v)
static int conv_weight(sparkcodec_handle* h, const std::string& prefix, std::initializer_list<int64_t> shape,
                       std::vector<float>* w) {
  const HostTensor* t;
  if (h->host.count(prefix + ".weight")) {
    SC_TRY(get(h, prefix + ".weight", &t, shape));
    *w = t->data;
    return 0;
  }
  const HostTensor *v, *g;
  SC_TRY(get(h, prefix + ".weight_v", &v, shape));
  SC_TRY(get(h, prefix + ".weight_g", &g));
  const int64_t d0 = v->shape[0];
  if (g->numel() != d0) {
    set_error("tensor '%s.weight_g' must have %lld elements", prefix.c_str(), (long long)d0);
    return SPARKCODEC_EINVAL;
  }
  fold_weight_norm(v->data.data(), g->data.data(), d0, v->numel() / d0, *w);
  return 0;
}

static int upload_snake(sparkcodec_handle* h, const std::string& key, int c, int replicate, SnakeParams* sp) {
  const HostTensor* t;
  SC_TRY(get(h, key, &t, {1, c, 1}));
  std::vector<float> a((size_t)c * replicate), inv((size_t)c * replicate);
  for (int r = 0; r < replicate; ++r)
    for (int i = 0; i < c; ++i) {
      a[(size_t)r * c + i] = t->data[i];
      inv[(size_t)r * c + i] = 1.0f / (t->data[i] + 1e-9f);   // (alpha + 1e-9).reciprocal(), layers.py:37
    }
  SC_TRY(upload(h, a, &sp->alpha));
  SC_TRY(upload(h, inv, &sp->inv));
  return 0;
}

static int build_conv1d(sparkcodec_handle* h, const std::string& prefix, int c_out, int c_in, int k, int dil,
                        bool linear_shape, const float* row_scale, GemmWeights* g) {
  std::vector<float> w;
  if (linear_shape) SC_TRY(conv_weight(h, prefix, {c_out, c_in}, &w));
  else SC_TRY(conv_weight(h, prefix, {c_out, c_in, k}, &w));
  const HostTensor* b;
  SC_TRY(get(h, prefix + ".bias", &b, {c_out}));
  PackedGemm pk;
  pack_conv1d(w.data(), c_out, c_in, k, dil, b->data.data(), row_scale, pk);
  return upload_gemm(h, pk, g);
}

static int build_backbone(sparkcodec_handle* h, const std::string& prefix, int layers, bool ada, float post_scale,
                          Backbone* bb, int c_in = 0) {
  const int C = h->cfg.vocos_dim, H = h->cfg.vocos_intermediate_dim;
  bb->ada = ada;
  SC_TRY(build_conv1d(h, prefix + ".embed", C, c_in ? c_in : C, 7, 1, false, nullptr, &bb->embed));
  const HostTensor *t, *u;
  if (!ada) {
    SC_TRY(get(h, prefix + ".norm.weight", &t, {C}));
    SC_TRY(get(h, prefix + ".norm.bias", &u, {C}));
    SC_TRY(upload(h, t->data, &bb->norm_w));
    SC_TRY(upload(h, u->data, &bb->norm_b));
  }
  bb->blocks.resize(layers);
  for (int i = 0; i < layers; ++i) {
    const std::string p = prefix + ".convnext." + std::to_string(i);
    ConvNeXt& blk = bb->blocks[i];
    SC_TRY(get(h, p + ".dwconv.weight", &t, {C, 1, 7}));
    std::vector<float> dw((size_t)7 * C);
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < 7; ++j) dw[(size_t)j * C + c] = t->data[(size_t)c * 7 + j];
    SC_TRY(upload(h, dw, &blk.dw_w));
    SC_TRY(get(h, p + ".dwconv.bias", &t, {C}));
    SC_TRY(upload(h, t->data, &blk.dw_b));
    if (!ada) {
      SC_TRY(get(h, p + ".norm.weight", &t, {C}));
      SC_TRY(get(h, p + ".norm.bias", &u, {C}));
      SC_TRY(upload(h, t->data, &blk.ln_w));
      SC_TRY(upload(h, u->data, &blk.ln_b));
    }
    SC_TRY(build_conv1d(h, p + ".pwconv1", H, C, 1, 1, true, nullptr, &blk.pw1));
    const HostTensor* gamma;
    SC_TRY(get(h, p + ".gamma", &gamma, {C}));
    SC_TRY(build_conv1d(h, p + ".pwconv2", C, H, 1, 1, true, gamma->data.data(), &blk.pw2));
  }
  SC_TRY(get(h, prefix + ".final_layer_norm.weight", &t, {C}));
  SC_TRY(get(h, prefix + ".final_layer_norm.bias", &u, {C}));
  std::vector<float> fw(t->data), fb(u->data);
  for (auto& v : fw) v *= post_scale;
  for (auto& v : fb) v *= post_scale;
  SC_TRY(upload(h, fw, &bb->final_w));
  SC_TRY(upload(h, fb, &bb->final_b));
  return 0;
}

// ---- speaker half of tokenize (bicodec.py:162-167), optional --------------------------------------------------
static int spk_conv(sparkcodec_handle* h, const std::string& conv, const std::string& bn, SpkConv* out) {
  const HostTensor *w, *b;
  SC_TRY(get(h, conv + ".weight", &w));
  if (w->shape.size() != 3) { set_error("tensor '%s.weight' must be (C_out, C_in, k)", conv.c_str()); return SPARKCODEC_EINVAL; }
  out->c_out = (int)w->shape[0]; out->c_in = (int)w->shape[1]; out->k = (int)w->shape[2];
  SC_TRY(get(h, conv + ".bias", &b, {out->c_out}));
  // (C_out, C_in, k) -> [n][j * C_in + c]: the layout the fp32 conv kernel reads
  std::vector<float> wt((size_t)out->c_out * out->k * out->c_in);
  for (int n = 0; n < out->c_out; ++n)
    for (int c = 0; c < out->c_in; ++c)
      for (int j = 0; j < out->k; ++j)
        wt[((size_t)n * out->k + j) * out->c_in + c] = w->data[((size_t)n * out->c_in + c) * out->k + j];
  SC_TRY(upload(h, wt, &out->w));
  SC_TRY(upload(h, b->data, &out->b));
  if (!bn.empty()) {   // BatchNorm1d(eval): y = (x - mean) / sqrt(var + 1e-5) * weight + bias  (ecapa_tdnn.py:63,113)
    const HostTensor *g, *be, *mu, *var;
    SC_TRY(get(h, bn + ".weight", &g, {out->c_out}));
    SC_TRY(get(h, bn + ".bias", &be, {out->c_out}));
    SC_TRY(get(h, bn + ".running_mean", &mu, {out->c_out}));
    SC_TRY(get(h, bn + ".running_var", &var, {out->c_out}));
    std::vector<float> sc(out->c_out), sh(out->c_out);
    for (int n = 0; n < out->c_out; ++n) {
      sc[n] = g->data[n] / std::sqrt(var->data[n] + 1e-5f);
      sh[n] = be->data[n] - mu->data[n] * sc[n];
    }
    SC_TRY(upload(h, sc, &out->scale));
    SC_TRY(upload(h, sh, &out->shift));
  }
  return 0;
}

static double hz_to_mel_slaney(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_hz / f_sp + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz_slaney(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, logstep = std::log(6.4) / 27.0, min_log_mel = min_log_hz / f_sp;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static int build_speaker(sparkcodec_handle* h) {
  SpeakerModel& S = h->spk;
  const sparkcodec_config& c = h->cfg;
  const std::string e = "speaker_encoder.speaker_encoder", ps = "speaker_encoder.perceiver_sampler";
  const HostTensor *t, *u;
  // ---- mel front end: windowed DFT rows and the slaney filterbank (torchaudio melscale_fbanks, norm = "slaney") ----
  if (S.win > S.n_fft || S.hop < 1 || S.n_mels < 1) { set_error("bad mel parameters"); return SPARKCODEC_EINVAL; }
  S.bins = S.n_fft / 2 + 1;
  {
    const double pi = 3.14159265358979323846;
    const int pad = (S.n_fft - S.win) / 2;      // torch.stft centres the window inside n_fft
    std::vector<float> dft((size_t)2 * S.bins * S.win);
    for (int k = 0; k < S.bins; ++k)
      for (int i = 0; i < S.win; ++i) {
        const double wv = 0.5 - 0.5 * std::cos(2.0 * pi * i / S.win);          // periodic Hann
        const double ang = 2.0 * pi * (double)k * (double)(i + pad) / S.n_fft;
        dft[(size_t)k * S.win + i] = (float)(wv * std::cos(ang));
        dft[(size_t)(S.bins + k) * S.win + i] = (float)(-wv * std::sin(ang));
      }
    SC_TRY(upload(h, dft, &S.dft));
    const double f_max = S.fmax < 0 ? S.sample_rate / 2 : S.fmax;
    const double m_lo = hz_to_mel_slaney(S.fmin), m_hi = hz_to_mel_slaney(f_max);
    std::vector<double> f_pts(S.n_mels + 2);
    for (int i = 0; i < S.n_mels + 2; ++i) f_pts[i] = mel_to_hz_slaney(m_lo + (m_hi - m_lo) * i / (S.n_mels + 1));
    std::vector<float> fb((size_t)S.n_mels * S.bins);
    for (int m = 0; m < S.n_mels; ++m) {
      const double enorm = 2.0 / (f_pts[m + 2] - f_pts[m]);
      for (int k = 0; k < S.bins; ++k) {
        const double f = (double)(S.sample_rate / 2) * k / (S.bins - 1);
        const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]), up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
        fb[(size_t)m * S.bins + k] = (float)(std::max(0.0, std::min(down, up)) * enorm);
      }
    }
    SC_TRY(upload(h, fb, &S.fb));
  }
  // ---- ECAPA-TDNN trunk (ecapa_tdnn.py:152-195) ----
  SC_TRY(spk_conv(h, e + ".layer1.conv", e + ".layer1.bn", &S.layer1));
  if (S.layer1.c_in != S.n_mels) { set_error("ECAPA layer1 expects %d mel bins, mel parameters give %d", S.layer1.c_in, S.n_mels); return SPARKCODEC_EINVAL; }
  S.channels = S.layer1.c_out;
  const int dil[3] = {2, 3, 4};
  S.blocks.resize(3);
  for (int i = 0; i < 3; ++i) {
    SpkRes2Block& B = S.blocks[i];
    const std::string p = e + ".layer" + std::to_string(i + 2) + ".se_res2block";
    B.dilation = dil[i];
    SC_TRY(spk_conv(h, p + ".0.conv", p + ".0.bn", &B.c0));
    for (int j = 0; h->host.count(p + ".1.convs." + std::to_string(j) + ".weight"); ++j) {
      B.convs.emplace_back();
      SC_TRY(spk_conv(h, p + ".1.convs." + std::to_string(j), p + ".1.bns." + std::to_string(j), &B.convs.back()));
    }
    if (B.convs.empty() || B.convs[0].k != 3) { set_error("unexpected Res2 block layout in '%s'", p.c_str()); return SPARKCODEC_EINVAL; }
    S.width = B.convs[0].c_out;
    if ((int)(B.convs.size() + 1) * S.width != S.channels) { set_error("Res2 widths do not tile %d channels", S.channels); return SPARKCODEC_EINVAL; }
    SC_TRY(spk_conv(h, p + ".2.conv", p + ".2.bn", &B.c2));
    SC_TRY(get(h, p + ".3.linear1.weight", &t));
    B.se_dim = (int)t->shape[0];
    SC_TRY(upload(h, t->data, &B.se_w1));
    SC_TRY(get(h, p + ".3.linear1.bias", &t, {B.se_dim}));
    SC_TRY(upload(h, t->data, &B.se_b1));
    SC_TRY(get(h, p + ".3.linear2.weight", &t, {S.channels, B.se_dim}));
    SC_TRY(upload(h, t->data, &B.se_w2));
    SC_TRY(get(h, p + ".3.linear2.bias", &t, {S.channels}));
    SC_TRY(upload(h, t->data, &B.se_b2));
  }
  SC_TRY(spk_conv(h, e + ".conv", "", &S.conv_cat));
  if (S.conv_cat.c_in != 3 * S.channels || S.conv_cat.k != 1) { set_error("ECAPA cat conv shape"); return SPARKCODEC_EINVAL; }
  // ---- perceiver resampler (perceiver_encoder.py:297-350) ----
  S.dim = c.latent_dim; S.n_lat = c.token_num;
  SC_TRY(get(h, ps + ".latents", &t, {S.n_lat, S.dim}));
  SC_TRY(upload(h, t->data, &S.latents));
  SC_TRY(get(h, ps + ".proj_context.weight", &t, {S.dim, S.conv_cat.c_out}));
  SC_TRY(upload(h, t->data, &S.pc_w));
  SC_TRY(get(h, ps + ".proj_context.bias", &t, {S.dim}));
  SC_TRY(upload(h, t->data, &S.pc_b));
  for (int i = 0; h->host.count(ps + ".layers." + std::to_string(i) + ".0.to_q.weight"); ++i) {
    const std::string a = ps + ".layers." + std::to_string(i) + ".0", f = ps + ".layers." + std::to_string(i) + ".1";
    S.layers.emplace_back();
    SpkPercLayer& L = S.layers.back();
    const int inner = S.heads * S.dim_head;
    SC_TRY(get(h, a + ".to_q.weight", &t, {inner, S.dim}));
    SC_TRY(upload(h, t->data, &L.wq));
    SC_TRY(get(h, a + ".to_kv.weight", &t, {2 * inner, S.dim}));
    SC_TRY(upload(h, t->data, &L.wkv));
    SC_TRY(get(h, a + ".to_out.weight", &t, {S.dim, inner}));
    SC_TRY(upload(h, t->data, &L.wo));
    SC_TRY(get(h, f + ".0.weight", &t));
    S.ff_inner = (int)t->shape[0] / 2;
    SC_TRY(upload(h, t->data, &L.w_ff0));
    SC_TRY(get(h, f + ".0.bias", &u, {2 * S.ff_inner}));
    SC_TRY(upload(h, u->data, &L.b_ff0));
    SC_TRY(get(h, f + ".2.weight", &t, {S.dim, S.ff_inner}));
    SC_TRY(upload(h, t->data, &L.w_ff2));
    SC_TRY(get(h, f + ".2.bias", &t, {S.dim}));
    SC_TRY(upload(h, t->data, &L.b_ff2));
  }
  SC_TRY(get(h, ps + ".norm.gamma", &t, {S.dim}));
  SC_TRY(upload(h, t->data, &S.norm_gamma));
  SC_TRY(get(h, "speaker_encoder.quantizer.project_in.weight", &t, {c.fsq_num_levels, S.dim}));
  SC_TRY(upload(h, t->data, &S.fsq_win));
  SC_TRY(get(h, "speaker_encoder.quantizer.project_in.bias", &t, {c.fsq_num_levels}));
  SC_TRY(upload(h, t->data, &S.fsq_bin));
  S.present = true;
  return 0;
}

static int do_finalize(sparkcodec_handle* h) {
  const sparkcodec_config& c = h->cfg;
  const int D = c.d_model, C = c.vocos_dim;
  const HostTensor *t, *u;

  void* ef;
  SC_TRY(dev_alloc(h, 4 * sizeof(int), &ef));
  h->err_flag = static_cast<int*>(ef);
  SC_CUDA(cudaMemset(h->err_flag, 0, 4 * sizeof(int)));

  // ---- quantizer: codebook, out_project, and the folded (x3 . linear_pre . out_project) map ----
  SC_TRY(get(h, "quantizer.codebook.weight", &t, {c.codebook_size, c.codebook_dim}));
  SC_TRY(upload(h, t->data, &h->codebook));
  std::vector<float> wout;
  SC_TRY(conv_weight(h, "quantizer.out_project", {D, c.codebook_dim, 1}, &wout));
  const HostTensor* bout;
  SC_TRY(get(h, "quantizer.out_project.bias", &bout, {D}));
  SC_TRY(upload(h, wout, &h->vq_w));
  SC_TRY(upload(h, bout->data, &h->vq_b));
  SC_TRY(get(h, "prenet.linear_pre.weight", &t, {C, D}));
  SC_TRY(get(h, "prenet.linear_pre.bias", &u, {C}));
  {
    // SamplingBlock with both ratios 1 returns 3*x (samper.py:79-100): one per downsample stage
    const double s3 = c.num_downsample > 0 ? 3.0 : 1.0;
    std::vector<float> mat((size_t)C * c.codebook_dim), vec(C);
    for (int o = 0; o < C; ++o) {
      for (int j = 0; j < c.codebook_dim; ++j) {
        double a = 0.0;
        for (int d = 0; d < D; ++d) a += (double)t->data[(size_t)o * D + d] * (double)wout[(size_t)d * c.codebook_dim + j];
        mat[(size_t)o * c.codebook_dim + j] = (float)(s3 * a);
      }
      double a = u->data[o];
      for (int d = 0; d < D; ++d) a += (double)t->data[(size_t)o * D + d] * (double)bout->data[d];
      vec[o] = (float)(s3 * a);
    }
    SC_TRY(upload(h, mat, &h->vq_mat));
    SC_TRY(upload(h, vec, &h->vq_vec));
  }

  // ---- speaker: FSQ project_out, project, stacked AdaLN scale/shift Linears ----
  {
    std::vector<int> lv(c.fsq_levels, c.fsq_levels + c.fsq_num_levels);
    SC_TRY(upload(h, lv, &h->fsq_levels));
  }
  SC_TRY(get(h, "speaker_encoder.quantizer.project_out.weight", &t, {c.latent_dim, c.fsq_num_levels}));
  SC_TRY(upload(h, t->data, &h->fsq_wpo));
  SC_TRY(get(h, "speaker_encoder.quantizer.project_out.bias", &t, {c.latent_dim}));
  SC_TRY(upload(h, t->data, &h->fsq_bpo));
  SC_TRY(get(h, "speaker_encoder.project.weight", &t, {D, (int64_t)c.latent_dim * c.token_num}));
  SC_TRY(upload(h, t->data, &h->spk_w));
  SC_TRY(get(h, "speaker_encoder.project.bias", &t, {D}));
  SC_TRY(upload(h, t->data, &h->spk_b));
  {
    // AdaLayerNorm i: rows [ (2i)*C, (2i+1)*C ) = scale, [ (2i+1)*C, (2i+2)*C ) = shift; i = 0 is the
    // backbone norm, i = 1.. the ConvNeXt norms (vocos.py:100-110, 293-306)
    h->n_ada = 1 + c.vocos_num_layers;
    std::vector<float> w((size_t)h->n_ada * 2 * C * D), b((size_t)h->n_ada * 2 * C);
    for (int i = 0; i < h->n_ada; ++i) {
      const std::string p = i == 0 ? std::string("prenet.vocos_backbone.norm")
                                   : "prenet.vocos_backbone.convnext." + std::to_string(i - 1) + ".norm";
      const char* names[2] = {".scale", ".shift"};
      for (int q = 0; q < 2; ++q) {
        SC_TRY(get(h, p + names[q] + ".weight", &t, {C, D}));
        SC_TRY(get(h, p + names[q] + ".bias", &u, {C}));
        memcpy(&w[((size_t)i * 2 + q) * C * D], t->data.data(), (size_t)C * D * sizeof(float));
        memcpy(&b[((size_t)i * 2 + q) * C], u->data.data(), (size_t)C * sizeof(float));
      }
    }
    SC_TRY(upload(h, w, &h->ada_w));
    SC_TRY(upload(h, b, &h->ada_b));
  }

  // ---- prenet backbones ----
  h->backbones.resize(c.num_downsample + 1);
  for (int i = 0; i < c.num_downsample; ++i)
    SC_TRY(build_backbone(h, "prenet.downsample." + std::to_string(i) + ".1", c.downsample_layers, false,
                          i + 1 < c.num_downsample ? 3.0f : 1.0f, &h->backbones[i]));
  SC_TRY(build_backbone(h, "prenet.vocos_backbone", c.vocos_num_layers, true, 1.0f, &h->backbones[c.num_downsample]));
  SC_TRY(build_conv1d(h, "prenet.linear", D, C, 1, 1, true, nullptr, &h->linear));

  // ---- wave generator ----
  const int ch = c.dec_channels;
  SC_TRY(build_conv1d(h, "decoder.model.0", ch, D, 7, 1, false, nullptr, &h->conv_in));
  h->ups.resize(c.num_upsample);
  int cin = ch;
  h->hop = 1;
  for (int i = 0; i < c.num_upsample; ++i) {
    UpBlock& ub = h->ups[i];
    const int k = c.kernel_sizes[i], s = c.rates[i], cout = ch >> (i + 1);
    if ((k - s) % 2 != 0 || k < s) {
      set_error("upsample %d: kernel %d / stride %d must satisfy (k - s) even and k >= s", i, k, s);
      return SPARKCODEC_EINVAL;
    }
    if (s > kMaxPhases || (k + s - 1) / s > kMaxTaps) {
      set_error("upsample %d: stride %d / kernel %d outside the supported polyphase range", i, s, k);
      return SPARKCODEC_EINVAL;
    }
    ub.stride = s; ub.c_in = cin; ub.c_out = cout;
    h->hop *= s;
    const std::string p = "decoder.model." + std::to_string(i + 1) + ".block";
    SnakeParams* s_in = i == 0 ? &h->s_conv_in : &h->ups[i - 1].ru[2].s_next;
    SC_TRY(upload_snake(h, p + ".0.alpha", cin, 1, s_in));
    std::vector<float> w;
    SC_TRY(conv_weight(h, p + ".1", {cin, cout, k}, &w));
    SC_TRY(get(h, p + ".1.bias", &t, {cout}));
    PackedGemm pk;
    pack_conv_transpose1d(w.data(), cin, cout, k, s, t->data.data(), pk);
    SC_TRY(upload_gemm(h, pk, &ub.convt));
    const int dil[3] = {1, 3, 9};
    for (int j = 0; j < 3; ++j) {
      const std::string q = p + "." + std::to_string(j + 2) + ".block";
      SnakeParams* s1 = j == 0 ? &ub.s_after : &ub.ru[j - 1].s_next;
      SC_TRY(upload_snake(h, q + ".0.alpha", cout, j == 0 ? s : 1, s1));
      SC_TRY(build_conv1d(h, q + ".1", cout, cout, 7, dil[j], false, nullptr, &ub.ru[j].c7));
      SC_TRY(upload_snake(h, q + ".2.alpha", cout, 1, &ub.ru[j].s_mid));
      SC_TRY(build_conv1d(h, q + ".3", cout, cout, 1, 1, false, nullptr, &ub.ru[j].c1));
    }
    cin = cout;
  }
  {
    const int n = c.num_upsample;
    SC_TRY(upload_snake(h, "decoder.model." + std::to_string(n + 1) + ".alpha", cin, 1, &h->s_head));
    std::vector<float> w;
    SC_TRY(conv_weight(h, "decoder.model." + std::to_string(n + 2), {1, cin, 7}, &w));
    std::vector<float> wt((size_t)7 * cin);
    for (int ci = 0; ci < cin; ++ci)
      for (int j = 0; j < 7; ++j) wt[(size_t)j * cin + ci] = w[(size_t)ci * 7 + j];
    SC_TRY(upload(h, wt, &h->head_w));
    SC_TRY(get(h, "decoder.model." + std::to_string(n + 2) + ".bias", &t, {1}));
    h->head_bias = t->data[0];
    h->head_c = cin;
  }
  // ---- encode side, semantic half (bicodec.py:151-169: encoder -> quantizer.tokenize); optional ----
  if (h->host.count("encoder.encoder.embed.weight")) {
    if (c.codebook_dim != 8) { set_error("semantic tokenize needs codebook_dim == 8"); return SPARKCODEC_EINVAL; }
    h->enc_backbones.resize(c.num_downsample + 1);
    // feat_encoder.py:66-77: VocosBackbone(input -> C, vocos_num_layers) then the down-sample stages; every
    // SamplingBlock (ratio 1) returns 3 x, folded into the preceding final LayerNorm like in the prenet
    SC_TRY(build_backbone(h, "encoder.encoder", c.vocos_num_layers, false, c.num_downsample > 0 ? 3.0f : 1.0f,
                          &h->enc_backbones[0], D));
    for (int i = 0; i < c.num_downsample; ++i)
      SC_TRY(build_backbone(h, "encoder.downsample." + std::to_string(i) + ".1", c.downsample_layers, false,
                            i + 1 < c.num_downsample ? 3.0f : 1.0f, &h->enc_backbones[i + 1]));
    // z_e = W_in (W_p x + b_p) + b_in: project (Linear C -> D) and in_project (1x1 conv D -> 8) have nothing
    // between them (feat_encoder.py:87-90, factorized_vector_quantize.py:147-150)
    std::vector<float> win;
    SC_TRY(conv_weight(h, "quantizer.in_project", {c.codebook_dim, D, 1}, &win));
    const HostTensor *bin, *wp, *bp;
    SC_TRY(get(h, "quantizer.in_project.bias", &bin, {c.codebook_dim}));
    SC_TRY(get(h, "encoder.project.weight", &wp, {D, C}));
    SC_TRY(get(h, "encoder.project.bias", &bp, {D}));
    std::vector<float> mat((size_t)c.codebook_dim * C), vec(c.codebook_dim);
    for (int j = 0; j < c.codebook_dim; ++j) {
      for (int o = 0; o < C; ++o) {
        double a = 0.0;
        for (int d = 0; d < D; ++d) a += (double)win[(size_t)j * D + d] * (double)wp->data[(size_t)d * C + o];
        mat[(size_t)j * C + o] = (float)a;
      }
      double a = bin->data[j];
      for (int d = 0; d < D; ++d) a += (double)win[(size_t)j * D + d] * (double)bp->data[d];
      vec[j] = (float)a;
    }
    SC_TRY(upload(h, mat, &h->tok_mat));
    SC_TRY(upload(h, vec, &h->tok_vec));
    SC_TRY(get(h, "quantizer.codebook.weight", &t, {c.codebook_size, c.codebook_dim}));
    std::vector<float> cn(t->data.size()), csq(c.codebook_size);
    for (int k = 0; k < c.codebook_size; ++k) {
      float n2 = 0.f;
      for (int j = 0; j < c.codebook_dim; ++j) n2 += t->data[(size_t)k * c.codebook_dim + j] * t->data[(size_t)k * c.codebook_dim + j];
      const float inv = 1.0f / std::max(std::sqrt(n2), 1e-12f);
      float s2 = 0.f;
      for (int j = 0; j < c.codebook_dim; ++j) {
        const float v = t->data[(size_t)k * c.codebook_dim + j] * inv;
        cn[(size_t)k * c.codebook_dim + j] = v;
        s2 += v * v;
      }
      csq[k] = s2;
    }
    SC_TRY(upload(h, cn, &h->codes_n));
    SC_TRY(upload(h, csq, &h->codes_sq));
    h->has_encoder = true;
  }
  if (h->host.count("speaker_encoder.speaker_encoder.layer1.conv.weight")) SC_TRY(build_speaker(h));
  h->host.clear();
  h->finalized = true;
  return 0;
}

// ------------------------------------------------------------------------------------ workspace
struct Arena {
  char* base; size_t off = 0, cap;
  bool overflow = false;
  void* take(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;
    void* p = base ? base + off : nullptr;
    off += bytes;
    if (off > cap) overflow = true;
    return p;
  }
  float* f32(size_t n) { return static_cast<float*>(take(n * 4)); }
  OpBuf op(size_t n) {
    OpBuf o;
    o.hi = static_cast<__nv_bfloat16*>(take(n * 2));
    o.lo = static_cast<__nv_bfloat16*>(take(n * 2));
    return o;
  }
};

struct Workspace {
  float *d, *flat, *ada, *px, *py, *x;
  OpBuf pa, ph, w_in, w_c0, op1[2], op2;
};

static size_t max_cl(const sparkcodec_handle* h) {   // max over up-blocks of (channels * rows per frame)
  size_t m = 0, rate = 1;
  for (auto& ub : h->ups) {
    rate *= ub.stride;
    m = std::max(m, (size_t)ub.c_out * rate);
  }
  return m;
}

static void carve(const sparkcodec_handle* h, Arena& a, size_t B, size_t T, Workspace* w) {
  const sparkcodec_config& c = h->cfg;
  const size_t F = B * T, C = c.vocos_dim, D = c.d_model;
  w->d = a.f32(B * D);
  w->flat = a.f32(B * (size_t)c.latent_dim * c.token_num);
  w->ada = a.f32(B * (size_t)h->n_ada * 2 * C);
  w->px = a.f32(F * C);
  w->py = a.f32(F * C);
  w->pa = a.op(F * C);
  w->ph = a.op(F * c.vocos_intermediate_dim);
  w->w_in = a.op(F * D);
  w->w_c0 = a.op(F * c.dec_channels);
  const size_t m = max_cl(h);
  w->x = a.f32(F * m);
  w->op1[0] = a.op(F * m);
  w->op1[1] = a.op(F * m);
  w->op2 = a.op(F * m);
}

static size_t workspace_needed(const sparkcodec_handle* h, size_t B, size_t T) {
  Arena a{nullptr, 0, ~(size_t)0};
  Workspace w;
  carve(h, a, B, T, &w);
  return a.off + 1024;
}

// ------------------------------------------------------------------------------------ schedule
struct TapReq {
  const char* name = nullptr;
  float* out = nullptr;
  size_t cap = 0;
  int64_t* shape = nullptr;
  size_t b0 = 0;   // first utterance of the current pass
  bool hit = false;
};

static OpBuf mode_op(OpBuf o, int prec) {
  if (!is_split(prec)) o.lo = nullptr;
  o.fmt = op_fmt_for(prec);
  return o;
}

struct Pass {
  sparkcodec_handle* h;
  int B, T, prec;
  cudaStream_t st;
  TapReq* tap;

  // ---- optional per-launch CUDA-event timing (profile mode only) ----
  int prof_begin(const std::string& name, double flops, double bytes) {
    if (!h->profile) return 0;
    sparkcodec_handle::ProfRec r{name, flops, bytes, nullptr, nullptr};
    SC_CUDA(cudaEventCreate(&r.e0));
    SC_CUDA(cudaEventCreate(&r.e1));
    SC_CUDA(cudaEventRecord(r.e0, st));
    h->prof.push_back(r);
    return 0;
  }
  int prof_end() {
    if (!h->profile) return 0;
    SC_CUDA(cudaEventRecord(h->prof.back().e1, st));
    return 0;
  }
  int gemm(const GemmWeights& w, const OpBuf& a, int L, const Epilogue& ep) {
    if (h->profile) {
      double macs = 0;
      int tmax = 0;
      for (int r = 0; r < w.taps.n_phase; ++r) {
        macs += (double)w.taps.ntaps[r] * w.c_in * w.taps.cols_per_phase;
        tmax = std::max(tmax, w.taps.ntaps[r]);
      }
      const double planes = is_split(prec) ? 4.0 : 2.0;   // bytes per operand element
      double bytes = (double)B * L * w.c_in * planes + (double)w.n_total * w.kt * w.c_in * planes;
      if (ep.residual) bytes += (double)B * L * w.n_total * 4;
      if (ep.out_f32) bytes += (double)B * L * w.n_total * 4;
      if (ep.out_op.hi) bytes += (double)B * L * w.n_total * planes;
      char nm[160];
      snprintf(nm, sizeof(nm), "conv_gemm_tc cin=%d n=%d phases=%d taps=%d L=%d bn=%d act=%d res=%d", w.c_in, w.n_total,
               w.taps.n_phase, tmax, L, w.block_n, ep.act, ep.residual != nullptr ? 1 : 0);
      SC_TRY(prof_begin(nm, 2.0 * B * L * macs, bytes));
    }
    int rc = h->impl == SPARKCODEC_IMPL_SIMT ? launch_conv_gemm_simt(w, a, B, L, ep, prec, st)
                                             : launch_conv_gemm_tc(w, a, B, L, ep, prec, h->num_sms, st);
    if (rc) return rc;
    return prof_end();
  }
  bool want(const char* name) const { return tap && tap->name && strcmp(tap->name, name) == 0; }
  int tap_check(size_t rows, size_t ch) {
    const size_t n = (size_t)B * rows * ch;
    if ((tap->b0 * rows * ch + n) > tap->cap) {
      set_error("tap buffer too small for '%s' (%zu floats needed)", tap->name, tap->b0 * rows * ch + n);
      return SPARKCODEC_ENOMEM;
    }
    if (tap->shape) { tap->shape[0] = (int64_t)rows; tap->shape[1] = (int64_t)ch; }
    tap->hit = true;
    return 0;
  }
  int tap_f32(const char* name, const float* src, size_t rows, size_t ch) {
    if (!want(name)) return 0;
    SC_TRY(tap_check(rows, ch));
    SC_CUDA(cudaMemcpyAsync(tap->out + tap->b0 * rows * ch, src, (size_t)B * rows * ch * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  int tap_op(const char* name, const OpBuf& src, size_t rows, size_t ch) {
    if (!want(name)) return 0;
    SC_TRY(tap_check(rows, ch));
    const OpBuf s = mode_op(src, prec);
    return launch_merge(s, tap->out + tap->b0 * rows * ch, (size_t)B * rows * ch, st);
  }
};

// One VocosBackbone (vocos.py:324-335) over operand planes `in` (B, T, c_in): embed conv k7 -> norm -> ConvNeXt
// blocks -> final LayerNorm, written as operand planes (out_op) or fp32 (out_f32).  Shared by the prenet and
// the encoder; `ada` = per-utterance AdaLN scale/shift rows (stride ada_n) or null.
static int run_backbone(Pass& P, Backbone& bb, const std::string& name, Workspace& W, const OpBuf& in, const float* ada,
                        int ada_n, float* out_f32, const OpBuf& out_op) {
  const int B = P.B, T = P.T, prec = P.prec, C = P.h->cfg.vocos_dim;
  cudaStream_t st = P.st;
  NvtxRange range("sparkcodec." + name);
  const OpBuf pa = mode_op(W.pa, prec), ph = mode_op(W.ph, prec);
  Epilogue e;
  e.out_f32 = W.py;
  SC_TRY(P.gemm(bb.embed, in, T, e));                                             // embed conv k7 -> py
  const float* sc = ada ? ada : bb.norm_w;
  const float* sh = ada ? ada + C : bb.norm_b;
  SC_TRY(P.prof_begin("ln", 0, (double)B * T * C * 8));
  SC_TRY(launch_dwconv_ln(W.py, B, T, C, nullptr, nullptr, sc, sh, ada ? ada_n : 0, 1e-6f, W.px, OpBuf(), st));
  SC_TRY(P.prof_end());
  SC_TRY(P.tap_f32((name + ".norm").c_str(), W.px, T, C));
  for (size_t i = 0; i < bb.blocks.size(); ++i) {
    ConvNeXt& blk = bb.blocks[i];
    sc = ada ? ada + (size_t)(i + 1) * 2 * C : blk.ln_w;
    sh = ada ? ada + (size_t)(i + 1) * 2 * C + C : blk.ln_b;
    SC_TRY(P.prof_begin("dwconv_ln", 0, (double)B * T * C * (4 + (is_split(prec) ? 4 : 2))));
    SC_TRY(launch_dwconv_ln(W.px, B, T, C, blk.dw_w, blk.dw_b, sc, sh, ada ? ada_n : 0, 1e-6f, nullptr, pa, st));
    SC_TRY(P.prof_end());
    Epilogue e1;
    e1.act = ACT_GELU;
    e1.out_op = ph;
    SC_TRY(P.gemm(blk.pw1, pa, T, e1));                                            // 384 -> 2048, GELU
    Epilogue e2;
    e2.residual = W.px;
    e2.out_f32 = W.px;
    SC_TRY(P.gemm(blk.pw2, ph, T, e2));                                            // 2048 -> 384, gamma, + x
    SC_TRY(P.tap_f32((name + ".convnext." + std::to_string(i)).c_str(), W.px, T, C));
  }
  return launch_dwconv_ln(W.px, B, T, C, nullptr, nullptr, bb.final_w, bb.final_b, 0, 1e-6f, out_f32, out_op, st);
}

// Semantic half of BiCodec.tokenize (bicodec.py:151-169): feat (B, T, D) fp32 -> encoder -> nearest code index.
static int run_tokenize(Pass& P, const float* feat, Workspace& W, long long* idx_out, float* margin_out) {
  sparkcodec_handle* h = P.h;
  const sparkcodec_config& c = h->cfg;
  const int B = P.B, T = P.T, prec = P.prec;
  const OpBuf pa = mode_op(W.pa, prec), w_in = mode_op(W.w_in, prec);
  SC_TRY(launch_split(feat, w_in, (size_t)B * T * c.d_model, P.st));
  for (size_t bi = 0; bi < h->enc_backbones.size(); ++bi) {
    const bool last = bi + 1 == h->enc_backbones.size();
    const std::string name = bi == 0 ? std::string("encoder.encoder") : "encoder.downsample." + std::to_string(bi - 1) + ".1";
    SC_TRY(run_backbone(P, h->enc_backbones[bi], name, W, bi == 0 ? w_in : pa, nullptr, 0, last ? W.py : nullptr,
                        last ? OpBuf() : pa));
  }
  SC_TRY(P.prof_begin("vq_search", 2.0 * B * T * (double)c.codebook_size * c.codebook_dim, (double)B * T * (c.vocos_dim * 4 + 8)));
  SC_TRY(launch_vq_search(W.py, (size_t)B * T, c.vocos_dim, h->tok_mat, h->tok_vec, h->codes_n, h->codes_sq,
                          c.codebook_size, c.codebook_dim, idx_out, margin_out, P.st));
  return P.prof_end();
}

// One pass over B utterances of T frames.  x_in != null: skip the token stages + prenet and start the
// wave generator from x_in (B,T,D) fp32.  x_out != null: stop after prenet(+d) and write it as fp32.
// staged: the operand planes of x are already in W.w_in (sparkcodec_wavegen_stage): start at the wave generator.
static int run_pass(Pass& P, const void* sem, int sem_dt, const void* glob, int glob_dt, Workspace& W,
                    const float* x_in, float* x_out, float* wav_out, bool staged = false) {
  sparkcodec_handle* h = P.h;
  const sparkcodec_config& c = h->cfg;
  const int B = P.B, T = P.T, prec = P.prec;
  const int C = c.vocos_dim, D = c.d_model;
  cudaStream_t st = P.st;
  const OpBuf pa = mode_op(W.pa, prec), w_in = mode_op(W.w_in, prec), w_c0 = mode_op(W.w_c0, prec),
              op2 = mode_op(W.op2, prec);
  const OpBuf op1[2] = {mode_op(W.op1[0], prec), mode_op(W.op1[1], prec)};

  if (staged) {
    // nothing to do: W.w_in was filled by sparkcodec_wavegen_stage
  } else if (!x_in) {
    NvtxRange range_tok("sparkcodec.token_stages+prenet");
    // ---- speaker tokens -> d_vector, AdaLN scale/shift for all 1 + vocos_num_layers norms ----
    SC_TRY(launch_fsq_project(glob, glob_dt, B, c.token_num, c.fsq_num_levels, h->fsq_levels, h->fsq_wpo, h->fsq_bpo,
                              c.latent_dim, W.flat, h->err_flag, st));
    if (P.want("fsq_codes")) {   // the index stage by itself: (B, token_num, n_levels) level codes, bit-exact
      SC_TRY(P.tap_check(c.token_num, c.fsq_num_levels));
      SC_TRY(launch_fsq_codes(glob, glob_dt, B * c.token_num, c.fsq_num_levels, h->fsq_levels,
                              P.tap->out + P.tap->b0 * (size_t)c.token_num * c.fsq_num_levels, st));
    }
    SC_TRY(launch_small_linear(W.flat, h->spk_w, h->spk_b, W.d, B, c.latent_dim * c.token_num, D, st));
    SC_TRY(P.tap_f32("d_vector", W.d, 1, D));
    const int ada_n = h->n_ada * 2 * C;
    SC_TRY(launch_small_linear(W.d, h->ada_w, h->ada_b, W.ada, B, D, ada_n, st));
    // ---- semantic tokens -> 3 * linear_pre(out_project(codebook[idx])) as operand planes ----
    if (P.want("codebook_rows")) {   // the index stage by itself: (B, T, codebook_dim) gathered rows, bit-exact
      SC_TRY(P.tap_check(T, c.codebook_dim));
      SC_TRY(launch_vq_rows(sem, sem_dt, B * T, c.codebook_size, c.codebook_dim, h->codebook,
                            P.tap->out + P.tap->b0 * (size_t)T * c.codebook_dim, st));
    }
    if (P.want("z_q")) {
      SC_TRY(P.tap_check(T, D));
      SC_TRY(launch_vq_zq(sem, sem_dt, B * T, c.codebook_size, c.codebook_dim, h->codebook, h->vq_w, h->vq_b, D,
                          P.tap->out + P.tap->b0 * (size_t)T * D, st));
    }
    SC_TRY(launch_vq_embed(sem, sem_dt, B, T, 0, T, c.codebook_size, c.codebook_dim, h->codebook, h->vq_mat,
                           h->vq_vec, C, pa, h->err_flag, st));
    // ---- prenet: 2 x VocosBackbone(2 layers) + VocosBackbone(12 layers, AdaLN) ----
    for (size_t bi = 0; bi < h->backbones.size(); ++bi) {
      Backbone& bb = h->backbones[bi];
      const std::string name = bb.ada ? std::string("prenet.vocos_backbone")
                                      : "prenet.downsample." + std::to_string(bi) + ".1";
      SC_TRY(run_backbone(P, bb, name, W, pa, bb.ada ? W.ada : nullptr, ada_n, nullptr, pa));
      // oracle tap names; NOTE the x3 of the following SamplingBlock is already folded in for downsample.0
      const std::string out_name = bb.ada ? std::string("prenet.vocos_backbone") : "prenet.downsample." + std::to_string(bi);
      SC_TRY(P.tap_op(out_name.c_str(), pa, T, C));
    }
    // ---- linear 384 -> 1024, + d_vector broadcast over time (bicodec.py:186) ----
    Epilogue el;
    el.rowbias = W.d;
    if (x_out) el.out_f32 = x_out; else el.out_op = w_in;
    if (P.want("prenet_plus_d") && !x_out) el.out_f32 = W.x;   // W.x is free until the first up-block
    SC_TRY(P.gemm(h->linear, pa, T, el));
    if (!x_out) SC_TRY(P.tap_f32("prenet_plus_d", W.x, T, D));
    if (x_out) return 0;
  } else {
    SC_TRY(launch_split(x_in, w_in, (size_t)B * T * D, st));
  }

  // ---- WaveGenerator ----
  NvtxRange range_wg("sparkcodec.wavegen");
  {
    NvtxRange range("sparkcodec.decoder.model.0");
    Epilogue e;
    e.act = ACT_SNAKE; e.alpha = h->s_conv_in.alpha; e.inv_alpha = h->s_conv_in.inv;
    e.out_op = w_c0;
    if (P.want("decoder.model.0")) e.out_f32 = W.x;
    SC_TRY(P.gemm(h->conv_in, w_in, T, e));
    SC_TRY(P.tap_f32("decoder.model.0", W.x, T, c.dec_channels));
  }
  OpBuf cur = w_c0;   // snake'd operand feeding the next transposed conv
  int L = T, pp = 0;
  for (size_t i = 0; i < h->ups.size(); ++i) {
    UpBlock& ub = h->ups[i];
    const std::string name = "decoder.model." + std::to_string(i + 1) + ".block";
    NvtxRange range("sparkcodec." + name);
    {
      if (cur.hi == op1[pp].hi) pp ^= 1;   // the transposed conv must not write over its own input
      Epilogue e;
      e.out_f32 = W.x;
      e.act = ACT_SNAKE; e.alpha = ub.s_after.alpha; e.inv_alpha = ub.s_after.inv;
      e.out_op = op1[pp];
      SC_TRY(P.gemm(ub.convt, cur, L, e));   // rows L, n_total = s*C_out  ==  (B, s*L, C_out)
    }
    L *= ub.stride;
    SC_TRY(P.tap_f32((name + ".1").c_str(), W.x, L, ub.c_out));
    for (int j = 0; j < 3; ++j) {
      ResUnit& ru = ub.ru[j];
      const bool last_unit = (i + 1 == h->ups.size()) && j == 2;
      int dil = 0;
      if (h->impl == SPARKCODEC_IMPL_TC && resunit_fusable(ru.c7, ru.c1, &dil)) {
        // narrow stages: the whole ResidualUnit is one kernel, `mid` never reaches HBM
        if (h->profile) {
          const double planes = is_split(prec) ? 4.0 : 2.0, C = ub.c_out;
          char nm[160];
          snprintf(nm, sizeof(nm), "resunit_fused c=%d dil=%d L=%d", ub.c_out, dil, L);
          SC_TRY(P.prof_begin(nm, 2.0 * B * L * 8.0 * C * C,
                              (double)B * L * C * (planes + 8.0 + (last_unit ? 0.0 : planes)) + 8.0 * C * C * planes));
        }
        SC_TRY(launch_resunit_fused(ru.c7, ru.c1, op1[pp], B, L, ru.s_mid.alpha, ru.s_mid.inv, W.x,
                                    last_unit ? nullptr : ru.s_next.alpha, last_unit ? nullptr : ru.s_next.inv,
                                    last_unit ? OpBuf() : op1[pp ^ 1], prec, h->num_sms, st));
        SC_TRY(P.prof_end());
        pp ^= 1;
        SC_TRY(P.tap_f32((name + "." + std::to_string(j + 2)).c_str(), W.x, L, ub.c_out));
        continue;
      }
      Epilogue e7;
      e7.act = ACT_SNAKE; e7.alpha = ru.s_mid.alpha; e7.inv_alpha = ru.s_mid.inv;
      e7.out_op = op2;
      SC_TRY(P.gemm(ru.c7, op1[pp], L, e7));
      Epilogue e1;
      e1.residual = W.x;
      e1.out_f32 = W.x;
      const bool last = (i + 1 == h->ups.size()) && j == 2;
      if (!last) {
        e1.act = ACT_SNAKE; e1.alpha = ru.s_next.alpha; e1.inv_alpha = ru.s_next.inv;
        e1.out_op = op1[pp ^ 1];
      }
      SC_TRY(P.gemm(ru.c1, op2, L, e1));
      pp ^= 1;
      SC_TRY(P.tap_f32((name + "." + std::to_string(j + 2)).c_str(), W.x, L, ub.c_out));
    }
    cur = op1[pp];
  }
  NvtxRange range_head("sparkcodec.head");
  SC_TRY(P.prof_begin("head", 0, (double)B * L * (h->head_c * 4 + 4)));
  // range-reduced sine only next to the three-term split (as in the conv epilogues, gemm_params.cuh)
  SC_TRY(launch_head(W.x, B, L, h->head_c, h->s_head.alpha, h->s_head.inv, h->head_w, h->head_bias, wav_out,
                     terms_for(P.prec) == 3 ? 1 : 0, L, st));
  SC_TRY(P.prof_end());
  return 0;
}

// ---- speaker half of tokenize: workspace + schedule ---------------------------------------------------------
struct SpkWorkspace {
  float *frames, *spec, *mag, *mel, *a1, *y, *z, *u, *cat, *latent, *ctx, *kv, *lat, *lat2, *q, *att, *ffh, *ffg, *se_m, *se_h, *se_s;
};
static int spk_frames_of(const SpeakerModel& S, int n_samples) { return n_samples / S.hop + 1; }   // torch.stft(center=True)
static void spk_carve(const sparkcodec_handle* h, Arena& a, size_t B, size_t n_samples, SpkWorkspace* w) {
  const SpeakerModel& S = h->spk;
  const size_t T = spk_frames_of(S, (int)n_samples), R = B * T, C = S.channels, inner = (size_t)S.heads * S.dim_head;
  w->frames = a.f32(R * S.win);
  w->spec = a.f32(R * 2 * S.bins);
  w->mag = a.f32(R * S.bins);
  w->mel = a.f32(R * S.n_mels);
  w->a1 = a.f32(R * C);
  w->y = a.f32(R * C);
  w->z = a.f32(R * C);
  w->u = a.f32(R * C);
  w->cat = a.f32(R * 3 * C);
  w->latent = a.f32(R * S.conv_cat.c_out);
  w->ctx = a.f32(B * (S.n_lat + T) * S.dim);
  w->kv = a.f32(B * (S.n_lat + T) * 2 * inner);
  w->lat = a.f32(B * S.n_lat * S.dim);
  w->lat2 = a.f32(B * S.n_lat * S.dim);
  w->q = a.f32(B * S.n_lat * inner);
  w->att = a.f32(B * S.n_lat * inner);
  w->ffh = a.f32(B * S.n_lat * 2 * S.ff_inner);
  w->ffg = a.f32(B * S.n_lat * S.ff_inner);
  w->se_m = a.f32(B * C);
  w->se_h = a.f32(B * 128 + B * 1024);
  w->se_s = a.f32(B * C);
}
static size_t spk_workspace_needed(const sparkcodec_handle* h, size_t B, size_t n_samples) {
  Arena a{nullptr, 0, ~(size_t)0};
  SpkWorkspace w;
  spk_carve(h, a, B, n_samples, &w);
  return a.off + 1024;
}

static int spk_run_conv(const SpkConv& cv, const float* x, int ldx, const float* x2, int ldx2, int B, int rows, int dilation,
                        bool relu, float* y, int ldy, cudaStream_t st) {
  SpkGemm g;
  g.x = x; g.ldx = ldx; g.x2 = x2; g.ldx2 = ldx2;
  g.batch = B; g.rows = rows; g.K = cv.c_in;
  g.ntaps = cv.k;
  for (int j = 0; j < cv.k; ++j) g.shift[j] = (j - (cv.k - 1) / 2) * dilation;   // "same" padding = dilation * (k - 1) / 2
  g.w = cv.w; g.ldw = cv.k * cv.c_in; g.bias = cv.b;
  g.relu = relu ? 1 : 0; g.scale = cv.scale; g.shift_v = cv.shift;
  g.y = y; g.ldy = ldy; g.N = cv.c_out;
  return launch_spk_gemm(g, st);
}
static int spk_linear(const float* x, int ldx, int rows, int K, const float* w, const float* bias, const float* res, int ldres,
                      int act_relu, int act, float* y, int ldy, int N, cudaStream_t st) {
  SpkGemm g;
  g.x = x; g.ldx = ldx; g.batch = 1; g.rows = rows; g.K = K;
  g.w = w; g.ldw = K; g.bias = bias; g.relu = act_relu; g.act = act;
  g.res = res; g.ldres = ldres; g.y = y; g.ldy = ldy; g.N = N;
  return launch_spk_gemm(g, st);
}

// ref_wav (B, n) fp32 -> global tokens (B, token_num) int32.  Optional tap (test hook): name -> fp32 copy.
static int run_speaker(sparkcodec_handle* h, const float* wav, int B, int n, SpkWorkspace& W, int* tokens, float* margin,
                       TapReq* tap, cudaStream_t st) {
  const SpeakerModel& S = h->spk;
  const int T = spk_frames_of(S, n), R = B * T, C = S.channels, inner = S.heads * S.dim_head;
  auto tap_out = [&](const char* name, const float* src, size_t rows, size_t ch) -> int {
    if (!tap || !tap->name || strcmp(tap->name, name) != 0) return 0;
    if ((size_t)B * rows * ch > tap->cap) { set_error("tap buffer too small for '%s'", name); return SPARKCODEC_ENOMEM; }
    if (tap->shape) { tap->shape[0] = (int64_t)rows; tap->shape[1] = (int64_t)ch; }
    tap->hit = true;
    SC_CUDA(cudaMemcpyAsync(tap->out, src, (size_t)B * rows * ch * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  NvtxRange range("sparkcodec.tokenize_speaker");
  // ---- mel spectrogram (bicodec.py:191-211 MelSpectrogram, power 1, slaney): frames -> |DFT| -> filterbank ----
  SC_TRY(launch_spk_frames(wav, B, n, T, S.hop, S.win, W.frames, st));
  SC_TRY(spk_linear(W.frames, S.win, R, S.win, S.dft, nullptr, nullptr, 0, 0, 0, W.spec, 2 * S.bins, 2 * S.bins, st));
  SC_TRY(launch_spk_magnitude(W.spec, (size_t)R, S.bins, W.mag, st));
  SC_TRY(spk_linear(W.mag, S.bins, R, S.bins, S.fb, nullptr, nullptr, 0, 0, 0, W.mel, S.n_mels, S.n_mels, st));
  SC_TRY(tap_out("mel", W.mel, T, S.n_mels));
  // ---- ECAPA-TDNN trunk -> latent (ecapa_tdnn.py:196-208) ----
  SC_TRY(spk_run_conv(S.layer1, W.mel, S.n_mels, nullptr, 0, B, T, 1, true, W.a1, C, st));
  const float* x_in = W.a1;
  int ld_in = C;
  for (size_t bi = 0; bi < S.blocks.size(); ++bi) {
    const SpkRes2Block& K = S.blocks[bi];
    const int wd = S.width, nums = (int)K.convs.size();
    SC_TRY(spk_run_conv(K.c0, x_in, ld_in, nullptr, 0, B, T, 1, true, W.y, C, st));
    for (int i = 0; i < nums; ++i)   // sp = conv(sp_prev + spx[i]) -> relu -> bn, written to slice i (ecapa_tdnn.py:66-79)
      SC_TRY(spk_run_conv(K.convs[i], W.y + (size_t)i * wd, C, i >= 1 ? W.z + (size_t)(i - 1) * wd : nullptr, C, B, T,
                          K.dilation, true, W.z + (size_t)i * wd, C, st));
    SC_TRY(launch_spk_copy_cols(W.y + (size_t)nums * wd, C, W.z + (size_t)nums * wd, C, (size_t)R, wd, st));
    SC_TRY(spk_run_conv(K.c2, W.z, C, nullptr, 0, B, T, 1, true, W.u, C, st));
    // SE_Connect (ecapa_tdnn.py:119-132) and the block's residual: out = x + u * sigmoid(W2 relu(W1 mean_t(u)))
    SC_TRY(launch_spk_mean_rows(W.u, B, T, C, C, W.se_m, st));
    SC_TRY(spk_linear(W.se_m, C, B, C, K.se_w1, K.se_b1, nullptr, 0, 1, 0, W.se_h, K.se_dim, K.se_dim, st));
    SC_TRY(spk_linear(W.se_h, K.se_dim, B, K.se_dim, K.se_w2, K.se_b2, nullptr, 0, 0, 1, W.se_s, C, C, st));
    float* out = W.cat + bi * (size_t)C;                       // torch.cat([out2, out3, out4], dim=1): slice bi of (B, T, 3C)
    SC_TRY(launch_spk_se_apply(x_in, ld_in, W.u, C, W.se_s, B, T, C, out, 3 * C, st));
    x_in = out;
    ld_in = 3 * C;
  }
  SC_TRY(spk_run_conv(S.conv_cat, W.cat, 3 * C, nullptr, 0, B, T, 1, true, W.latent, S.conv_cat.c_out, st));
  SC_TRY(tap_out("ecapa_latent", W.latent, T, S.conv_cat.c_out));
  // ---- perceiver resampler (perceiver_encoder.py:335-347) ----
  const int nk = S.n_lat + T, D = S.dim;
  // proj_context straight into rows [n_lat, n_lat + T) of every utterance's context block
  for (int b = 0; b < B; ++b)
    SC_TRY(spk_linear(W.latent + (size_t)b * T * S.conv_cat.c_out, S.conv_cat.c_out, T, S.conv_cat.c_out, S.pc_w, S.pc_b,
                      nullptr, 0, 0, 0, W.ctx + ((size_t)b * nk + S.n_lat) * D, D, D, st));
  SC_TRY(launch_spk_place_rows(S.latents, 0, B, S.n_lat, D, W.lat, S.n_lat, 0, st));   // latents repeated per utterance
  for (const SpkPercLayer& L : S.layers) {
    SC_TRY(launch_spk_place_rows(W.lat, (size_t)S.n_lat * D, B, S.n_lat, D, W.ctx, nk, 0, st));   // context = cat(latents, x)
    SC_TRY(spk_linear(W.lat, D, B * S.n_lat, D, L.wq, nullptr, nullptr, 0, 0, 0, W.q, inner, inner, st));
    SC_TRY(spk_linear(W.ctx, D, B * nk, D, L.wkv, nullptr, nullptr, 0, 0, 0, W.kv, 2 * inner, 2 * inner, st));
    SC_TRY(launch_spk_attention(W.q, W.kv, B, S.n_lat, nk, S.heads, S.dim_head, W.att, st));
    SC_TRY(spk_linear(W.att, inner, B * S.n_lat, inner, L.wo, nullptr, W.lat, D, 0, 0, W.lat2, D, D, st));      // + latents
    SC_TRY(spk_linear(W.lat2, D, B * S.n_lat, D, L.w_ff0, L.b_ff0, nullptr, 0, 0, 0, W.ffh, 2 * S.ff_inner, 2 * S.ff_inner, st));
    SC_TRY(launch_spk_geglu(W.ffh, (size_t)B * S.n_lat, S.ff_inner, W.ffg, st));
    SC_TRY(spk_linear(W.ffg, S.ff_inner, B * S.n_lat, S.ff_inner, L.w_ff2, L.b_ff2, W.lat2, D, 0, 0, W.lat, D, D, st));   // + latents
  }
  SC_TRY(launch_spk_rmsnorm(W.lat, S.norm_gamma, D, B * S.n_lat, W.lat2, st));
  SC_TRY(tap_out("perceiver", W.lat2, S.n_lat, D));
  // ---- FSQ quantise (residual_fsq.py:213-283 with one quantizer) ----
  return launch_spk_fsq_quantize(W.lat2, D, B * S.n_lat, S.fsq_win, S.fsq_bin, h->fsq_levels, h->cfg.fsq_num_levels, tokens, margin, st);
}

static int check_common(sparkcodec_handle* h, int batch, int frames, int precision) {
  if (!h) { set_error("null handle"); return SPARKCODEC_EINVAL; }
  if (!h->finalized) { set_error("sparkcodec_finalize has not been called"); return SPARKCODEC_ESTATE; }
  if (batch < 0 || frames < 0) { set_error("negative batch/frames"); return SPARKCODEC_EINVAL; }
  if (precision != SPARKCODEC_PREC_FP32 && precision != SPARKCODEC_PREC_BF16 && precision != SPARKCODEC_PREC_FP32X3) {
    set_error("unknown precision %d", precision);
    return SPARKCODEC_EINVAL;
  }
  return 0;
}

// Splits the batch into passes that fit the caller's workspace.
static int run_all(sparkcodec_handle* h, const void* sem, int sem_dt, const void* glob, int glob_dt, int batch,
                   int frames, int precision, void* ws, size_t ws_bytes, const float* x_in, float* x_out,
                   float* wav_out, TapReq* tap, cudaStream_t st) {
  SC_TRY(check_common(h, batch, frames, precision));
  if ((sem_dt != SPARKCODEC_I32 && sem_dt != SPARKCODEC_I64) || (glob_dt != SPARKCODEC_I32 && glob_dt != SPARKCODEC_I64)) {
    set_error("token dtype must be SPARKCODEC_I32 or SPARKCODEC_I64");
    return SPARKCODEC_EINVAL;
  }
  if (batch == 0 || frames == 0) return 0;
  int per_pass = batch;
  while (per_pass > 1 && workspace_needed(h, per_pass, frames) > ws_bytes) per_pass = (per_pass + 1) / 2;
  if (workspace_needed(h, per_pass, frames) > ws_bytes || !ws) {
    set_error("workspace of %zu bytes cannot hold even one utterance of %d frames (%zu needed)", ws_bytes, frames,
              workspace_needed(h, 1, frames));
    return SPARKCODEC_ENOMEM;
  }
  SC_ON_DEVICE(h->device);
  g_launch_counter = &h->launches;
  const sparkcodec_config& c = h->cfg;
  const size_t sem_sz = sem_dt == SPARKCODEC_I64 ? 8 : 4, glob_sz = glob_dt == SPARKCODEC_I64 ? 8 : 4;
  for (int b0 = 0; b0 < batch; b0 += per_pass) {
    const int B = std::min(per_pass, batch - b0);
    Arena a{static_cast<char*>(ws), 0, ws_bytes};
    Workspace W;
    carve(h, a, B, frames, &W);
    Pass P{h, B, frames, precision, st, tap};
    NvtxRange range_pass(x_out ? "sparkcodec.prenet_pass" : x_in ? "sparkcodec.wavegen_pass" : "sparkcodec.detokenize_pass");
    if (tap) tap->b0 = b0;
    const char* semp = sem ? static_cast<const char*>(sem) + (size_t)b0 * frames * sem_sz : nullptr;
    const char* globp = glob ? static_cast<const char*>(glob) + (size_t)b0 * c.token_num * glob_sz : nullptr;
    SC_TRY(run_pass(P, semp, sem_dt, globp, glob_dt, W,
                    x_in ? x_in + (size_t)b0 * frames * c.d_model : nullptr,
                    x_out ? x_out + (size_t)b0 * frames * c.d_model : nullptr,
                    wav_out ? wav_out + (size_t)b0 * frames * h->hop : nullptr));
  }
  if (tap && tap->name && !tap->hit) {
    set_error("unknown tap '%s'", tap->name);
    return SPARKCODEC_EINVAL;
  }
  return 0;
}

}  // namespace sparkcodec

// ================================================================================== extern "C"
extern "C" {

const char* sparkcodec_last_error(void) { return g_err; }
int sparkcodec_abi_version(void) { return SPARKCODEC_ABI_VERSION; }

int sparkcodec_create(const sparkcodec_config* cfg, int device, sparkcodec_handle** out) {
  if (!cfg || !out) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  if (cfg->fsq_num_levels < 1 || cfg->fsq_num_levels > 8 || cfg->num_upsample < 1 || cfg->num_upsample > 8 ||
      cfg->num_downsample < 0 || cfg->num_downsample > 4) {
    set_error("config out of range (fsq_num_levels %d, num_upsample %d, num_downsample %d)", cfg->fsq_num_levels,
              cfg->num_upsample, cfg->num_downsample);
    return SPARKCODEC_EINVAL;
  }
  if (cfg->vocos_dim % 128 || cfg->d_model % 64 || cfg->vocos_intermediate_dim % 64 ||
      (cfg->dec_channels >> cfg->num_upsample) % 32 || (cfg->dec_channels >> cfg->num_upsample) > 128) {
    set_error("unsupported widths: vocos_dim %% 128, d_model/intermediate %% 64, final decoder width in {32,64,96,128}");
    return SPARKCODEC_EINVAL;
  }
  int ndev = 0;
  SC_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return SPARKCODEC_EINVAL; }
  SC_ON_DEVICE(device);
  cudaDeviceProp prop;
  SC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("libsparkcodec is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
    return SPARKCODEC_ECUDA;
  }
  auto* h = new sparkcodec_handle();
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  *out = h;
  return 0;
}

int sparkcodec_destroy(sparkcodec_handle* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  h->prof.clear();
  for (void* p : h->allocs) cudaFree(p);
  if (g_launch_counter == &h->launches) g_launch_counter = nullptr;
  delete h;
  return 0;
}

int sparkcodec_set_tensor(sparkcodec_handle* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  if (!h || !key || !data || (!shape && ndim > 0) || ndim < 0 || ndim > 4) { set_error("bad argument"); return SPARKCODEC_EINVAL; }
  if (h->finalized) { set_error("weights are already finalized"); return SPARKCODEC_ESTATE; }
  const std::string k(key);
  static const char* used[] = {"quantizer.codebook.", "quantizer.out_project.", "speaker_encoder.quantizer.project_out.",
                               "speaker_encoder.project.", "prenet.", "decoder.", "encoder.", "quantizer.in_project.",
                               "speaker_encoder.speaker_encoder.layer", "speaker_encoder.speaker_encoder.conv.",
                               "speaker_encoder.perceiver_sampler.", "speaker_encoder.quantizer.project_in."};
  bool keep = false;
  for (const char* p : used) keep |= k.rfind(p, 0) == 0;
  if (!keep) return 0;   // encode-side / training-only tensor
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  if (n < 0 || n > (int64_t)1 << 31) { set_error("tensor '%s' too large", key); return SPARKCODEC_EINVAL; }
  t.data.assign(data, data + n);
  h->host[k] = std::move(t);
  return 0;
}

int sparkcodec_finalize(sparkcodec_handle* h) {
  if (!h) { set_error("null handle"); return SPARKCODEC_EINVAL; }
  if (h->finalized) return 0;
  SC_ON_DEVICE(h->device);
  return do_finalize(h);
}

int sparkcodec_workspace_bytes(sparkcodec_handle* h, int batch, int frames, size_t* bytes) {
  if (!bytes) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  SC_TRY(check_common(h, batch, frames, SPARKCODEC_PREC_FP32));
  *bytes = workspace_needed(h, (size_t)std::max(batch, 1), (size_t)std::max(frames, 1));
  return 0;
}

int sparkcodec_detokenize(sparkcodec_handle* h, const void* semantic, int sem_dtype, const void* global_tokens,
                          int glob_dtype, int batch, int frames, int precision, void* workspace,
                          size_t workspace_bytes, float* wav_out, void* stream) {
  if (batch > 0 && frames > 0 && (!semantic || !global_tokens || !wav_out)) { set_error("null tensor pointer"); return SPARKCODEC_EINVAL; }
  return run_all(h, semantic, sem_dtype, global_tokens, glob_dtype, batch, frames, precision, workspace,
                 workspace_bytes, nullptr, nullptr, wav_out, nullptr, static_cast<cudaStream_t>(stream));
}

int sparkcodec_prenet(sparkcodec_handle* h, const void* semantic, int sem_dtype, const void* global_tokens,
                      int glob_dtype, int batch, int frames, int precision, void* workspace, size_t workspace_bytes,
                      float* x_out, void* stream) {
  if (batch > 0 && frames > 0 && (!semantic || !global_tokens || !x_out)) { set_error("null tensor pointer"); return SPARKCODEC_EINVAL; }
  return run_all(h, semantic, sem_dtype, global_tokens, glob_dtype, batch, frames, precision, workspace,
                 workspace_bytes, nullptr, x_out, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int sparkcodec_wavegen(sparkcodec_handle* h, const float* x_in, int batch, int frames, int precision, void* workspace,
                       size_t workspace_bytes, float* wav_out, void* stream) {
  if (batch > 0 && frames > 0 && (!x_in || !wav_out)) { set_error("null tensor pointer"); return SPARKCODEC_EINVAL; }
  return run_all(h, nullptr, SPARKCODEC_I32, nullptr, SPARKCODEC_I32, batch, frames, precision, workspace,
                 workspace_bytes, x_in, nullptr, wav_out, nullptr, static_cast<cudaStream_t>(stream));
}

// ---- staged WaveGenerator input (time sharding with the halo exchange overlapped, include/sparkcodec.h) ----
static int staged_workspace(sparkcodec_handle* h, int batch, int frames_total, int precision, void* workspace,
                            size_t workspace_bytes, Workspace* W) {
  SC_TRY(check_common(h, batch, frames_total, precision));
  if (batch == 0 || frames_total == 0) { set_error("empty staged call"); return SPARKCODEC_EINVAL; }
  if (!workspace || workspace_needed(h, batch, frames_total) > workspace_bytes) {
    set_error("staged wavegen needs the whole batch in one pass: workspace of %zu bytes, %zu needed", workspace_bytes,
              workspace_needed(h, batch, frames_total));
    return SPARKCODEC_ENOMEM;
  }
  Arena a{static_cast<char*>(workspace), 0, workspace_bytes};
  carve(h, a, batch, frames_total, W);
  return 0;
}

int sparkcodec_wavegen_stage(sparkcodec_handle* h, const float* x_rows, int batch, int rows, int64_t src_batch_stride,
                             int frames_total, int row_offset, int precision, void* workspace, size_t workspace_bytes,
                             void* stream) {
  Workspace W;
  SC_TRY(staged_workspace(h, batch, frames_total, precision, workspace, workspace_bytes, &W));
  if (rows == 0) return 0;
  if (!x_rows || rows < 0 || row_offset < 0 || row_offset + rows > frames_total ||
      src_batch_stride < (int64_t)rows * h->cfg.d_model) {
    set_error("wavegen_stage: rows [%d, %d) outside the %d-frame window or bad stride", row_offset, row_offset + rows,
              frames_total);
    return SPARKCODEC_EINVAL;
  }
  SC_ON_DEVICE(h->device);
  g_launch_counter = &h->launches;
  return launch_split_rows(x_rows, (size_t)src_batch_stride, mode_op(W.w_in, precision), batch, rows, h->cfg.d_model,
                           frames_total, row_offset, static_cast<cudaStream_t>(stream));
}

int sparkcodec_wavegen_staged(sparkcodec_handle* h, int batch, int frames_total, int precision, void* workspace,
                              size_t workspace_bytes, float* wav_out, void* stream) {
  Workspace W;
  SC_TRY(staged_workspace(h, batch, frames_total, precision, workspace, workspace_bytes, &W));
  if (!wav_out) { set_error("null tensor pointer"); return SPARKCODEC_EINVAL; }
  SC_ON_DEVICE(h->device);
  g_launch_counter = &h->launches;
  Pass P{h, batch, frames_total, precision, static_cast<cudaStream_t>(stream), nullptr};
  NvtxRange range_pass("sparkcodec.wavegen_staged_pass");
  return run_pass(P, nullptr, SPARKCODEC_I32, nullptr, SPARKCODEC_I32, W, nullptr, nullptr, wav_out, /*staged=*/true);
}

int sparkcodec_tokenize_semantic(sparkcodec_handle* h, const float* feat, int batch, int frames, int precision,
                                 void* workspace, size_t workspace_bytes, int64_t* tokens_out, float* margin_out,
                                 void* stream) {
  SC_TRY(check_common(h, batch, frames, precision));
  if (!h->has_encoder) {
    set_error("the checkpoint had no encoder.* / quantizer.in_project.* tensors: semantic tokenize is unavailable");
    return SPARKCODEC_ESTATE;
  }
  if (batch == 0 || frames == 0) return 0;
  if (!feat || !tokens_out) { set_error("null tensor pointer"); return SPARKCODEC_EINVAL; }
  int per_pass = batch;
  while (per_pass > 1 && workspace_needed(h, per_pass, frames) > workspace_bytes) per_pass = (per_pass + 1) / 2;
  if (workspace_needed(h, per_pass, frames) > workspace_bytes || !workspace) {
    set_error("workspace of %zu bytes cannot hold even one utterance of %d frames (%zu needed)", workspace_bytes,
              frames, workspace_needed(h, 1, frames));
    return SPARKCODEC_ENOMEM;
  }
  SC_ON_DEVICE(h->device);
  g_launch_counter = &h->launches;
  for (int b0 = 0; b0 < batch; b0 += per_pass) {
    const int B = std::min(per_pass, batch - b0);
    Arena a{static_cast<char*>(workspace), 0, workspace_bytes};
    Workspace W;
    carve(h, a, B, frames, &W);
    Pass P{h, B, frames, precision, static_cast<cudaStream_t>(stream), nullptr};
    SC_TRY(run_tokenize(P, feat + (size_t)b0 * frames * h->cfg.d_model, W,
                        reinterpret_cast<long long*>(tokens_out) + (size_t)b0 * frames,
                        margin_out ? margin_out + (size_t)b0 * frames : nullptr));
  }
  return 0;
}

int sparkcodec_set_mel_params(sparkcodec_handle* h, int sample_rate, int n_fft, int win_length, int hop_length, int n_mels,
                              float f_min, float f_max) {
  if (!h) { set_error("null handle"); return SPARKCODEC_EINVAL; }
  if (h->finalized) { set_error("weights are already finalized"); return SPARKCODEC_ESTATE; }
  if (sample_rate < 1 || n_fft < 2 || win_length < 1 || win_length > n_fft || hop_length < 1 || n_mels < 1 || f_min < 0) {
    set_error("bad mel parameters");
    return SPARKCODEC_EINVAL;
  }
  SpeakerModel& S = h->spk;
  S.sample_rate = sample_rate; S.n_fft = n_fft; S.win = win_length; S.hop = hop_length; S.n_mels = n_mels;
  S.fmin = f_min; S.fmax = f_max;
  return 0;
}

static int speaker_ready(sparkcodec_handle* h, int batch, int n_samples) {
  if (!h) { set_error("null handle"); return SPARKCODEC_EINVAL; }
  if (!h->finalized) { set_error("sparkcodec_finalize has not been called"); return SPARKCODEC_ESTATE; }
  if (!h->spk.present) {
    set_error("the checkpoint had no speaker_encoder.speaker_encoder.* / perceiver_sampler.* tensors: speaker tokenize is unavailable");
    return SPARKCODEC_ESTATE;
  }
  if (batch < 0 || n_samples < 0) { set_error("negative batch / samples"); return SPARKCODEC_EINVAL; }
  if (batch > 0 && n_samples <= h->spk.win / 2) {
    set_error("reference clip of %d samples is shorter than the STFT padding (%d)", n_samples, h->spk.win / 2 + 1);
    return SPARKCODEC_EINVAL;
  }
  return 0;
}

int sparkcodec_speaker_workspace_bytes(sparkcodec_handle* h, int batch, int n_samples, size_t* bytes) {
  if (!bytes) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  SC_TRY(speaker_ready(h, batch, n_samples));
  *bytes = spk_workspace_needed(h, (size_t)std::max(batch, 1), (size_t)std::max(n_samples, 1));
  return 0;
}

static int tokenize_speaker_impl(sparkcodec_handle* h, const float* ref_wav, int batch, int n_samples, void* workspace,
                                 size_t workspace_bytes, int32_t* tokens_out, float* margin_out, TapReq* tap, void* stream) {
  SC_TRY(speaker_ready(h, batch, n_samples));
  if (batch == 0) return 0;
  if (!ref_wav || !tokens_out) { set_error("null tensor pointer"); return SPARKCODEC_EINVAL; }
  if (!workspace || spk_workspace_needed(h, batch, n_samples) > workspace_bytes) {
    set_error("speaker workspace of %zu bytes is too small (%zu needed)", workspace_bytes, spk_workspace_needed(h, batch, n_samples));
    return SPARKCODEC_ENOMEM;
  }
  SC_ON_DEVICE(h->device);
  g_launch_counter = &h->launches;
  Arena a{static_cast<char*>(workspace), 0, workspace_bytes};
  SpkWorkspace W;
  spk_carve(h, a, batch, n_samples, &W);
  SC_TRY(run_speaker(h, ref_wav, batch, n_samples, W, tokens_out, margin_out, tap, static_cast<cudaStream_t>(stream)));
  if (tap && tap->name && !tap->hit) { set_error("unknown tap '%s'", tap->name); return SPARKCODEC_EINVAL; }
  return 0;
}

int sparkcodec_tokenize_speaker(sparkcodec_handle* h, const float* ref_wav, int batch, int n_samples, void* workspace,
                                size_t workspace_bytes, int32_t* tokens_out, float* margin_out, void* stream) {
  return tokenize_speaker_impl(h, ref_wav, batch, n_samples, workspace, workspace_bytes, tokens_out, margin_out, nullptr, stream);
}

int sparkcodec_tokenize_speaker_tap(sparkcodec_handle* h, const float* ref_wav, int batch, int n_samples, void* workspace,
                                    size_t workspace_bytes, int32_t* tokens_out, const char* tap, float* tap_out,
                                    size_t tap_capacity, int64_t* tap_shape, void* stream) {
  if (!tap || !tap_out) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  TapReq t;
  t.name = tap; t.out = tap_out; t.cap = tap_capacity; t.shape = tap_shape;
  return tokenize_speaker_impl(h, ref_wav, batch, n_samples, workspace, workspace_bytes, tokens_out, nullptr, &t, stream);
}

int sparkcodec_halo_frames(sparkcodec_handle* h, int* prenet_halo, int* wavegen_halo) {
  if (!h || !prenet_halo || !wavegen_halo) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  if (!h->finalized) { set_error("sparkcodec_finalize has not been called"); return SPARKCODEC_ESTATE; }
  const sparkcodec_config& c = h->cfg;
  // prenet: every backbone = embed conv (3) + layers x depthwise conv (3 each), all at frame rate
  int pre = 0;
  for (int i = 0; i < c.num_downsample; ++i) pre += 3 + 3 * c.downsample_layers;
  pre += 3 + 3 * c.vocos_num_layers;
  // wave generator: conv_in (3 frames), then per up-block the transposed conv's reach in input rows
  // (max |tap shift|) and three residual units (k7, dilation 1/3/9 -> 39 rows) at the block's rate
  double halo_frames = 3.0, rate = 1.0;
  for (auto& ub : h->ups) {
    int reach = 0;
    for (int r = 0; r < ub.convt.taps.n_phase; ++r)
      for (int m = 0; m < ub.convt.taps.ntaps[r]; ++m) reach = std::max(reach, std::abs(ub.convt.taps.shift[r][m]));
    halo_frames += reach / rate;
    rate *= ub.stride;
    halo_frames += 3.0 * (1 + 3 + 9) / rate;
  }
  halo_frames += 3.0 / rate;   // head conv k7
  *prenet_halo = pre;
  *wavegen_halo = (int)std::ceil(halo_frames);
  return 0;
}

int sparkcodec_check_tokens(sparkcodec_handle* h, void* stream) {
  if (!h || !h->finalized) { set_error("handle not ready"); return SPARKCODEC_ESTATE; }
  SC_ON_DEVICE(h->device);
  int e[4];
  SC_CUDA(cudaMemcpyAsync(e, h->err_flag, sizeof(e), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  SC_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  if (e[0] == 0) return 0;
  SC_CUDA(cudaMemsetAsync(h->err_flag, 0, sizeof(e), static_cast<cudaStream_t>(stream)));
  const long long v = ((long long)e[3] << 32) | (unsigned int)e[2];
  set_error("%s token id %lld at flat position %d is out of range", e[0] == 1 ? "semantic" : "global", v, e[1]);
  return SPARKCODEC_EINDEX;
}

int sparkcodec_extract_codes(const void* token_ids, int id_dtype, int batch, int n_tokens, int64_t semantic_base,
                             int codebook_size, int64_t global_base, int global_size, int32_t* semantic_out,
                             int32_t* semantic_len, int32_t* global_out, int max_global, int32_t* global_len,
                             void* stream) {
  if (batch < 0 || n_tokens < 0 || max_global < 0 || codebook_size < 0 || global_size < 0) { set_error("negative size"); return SPARKCODEC_EINVAL; }
  if (id_dtype != SPARKCODEC_I32 && id_dtype != SPARKCODEC_I64) { set_error("token dtype must be SPARKCODEC_I32 or SPARKCODEC_I64"); return SPARKCODEC_EINVAL; }
  if (batch == 0) return 0;
  if (!token_ids || !semantic_out || !semantic_len || !global_out || !global_len) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  g_launch_counter = nullptr;
  return launch_extract_codes(token_ids, id_dtype, batch, n_tokens, semantic_base, codebook_size, global_base, global_size,
                              semantic_out, semantic_len, global_out, max_global, global_len,
                              static_cast<cudaStream_t>(stream));
}

int sparkcodec_set_impl(sparkcodec_handle* h, int impl) {
  if (!h || (impl != SPARKCODEC_IMPL_TC && impl != SPARKCODEC_IMPL_SIMT && impl != SPARKCODEC_IMPL_TC_UNFUSED)) {
    set_error("bad impl");
    return SPARKCODEC_EINVAL;
  }
  h->impl = impl;
  return 0;
}

int sparkcodec_detokenize_tap(sparkcodec_handle* h, const void* semantic, int sem_dtype, const void* global_tokens,
                              int glob_dtype, int batch, int frames, int precision, void* workspace,
                              size_t workspace_bytes, float* wav_out, const char* tap, float* tap_out,
                              size_t tap_capacity, int64_t* tap_shape, void* stream) {
  if (!semantic || !global_tokens || !wav_out || !tap || !tap_out) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  TapReq t;
  t.name = tap; t.out = tap_out; t.cap = tap_capacity; t.shape = tap_shape;
  return run_all(h, semantic, sem_dtype, global_tokens, glob_dtype, batch, frames, precision, workspace,
                 workspace_bytes, nullptr, nullptr, wav_out, &t, static_cast<cudaStream_t>(stream));
}

int sparkcodec_profile(sparkcodec_handle* h, int enable) {
  if (!h) { set_error("null handle"); return SPARKCODEC_EINVAL; }
  SC_ON_DEVICE(h->device);
  for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  h->prof.clear();
  h->profile = enable != 0;
  return 0;
}

int sparkcodec_profile_read(sparkcodec_handle* h, char* buf, size_t cap, size_t* needed) {
  if (!h || !needed) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  SC_ON_DEVICE(h->device);
  std::string out;
  for (auto& r : h->prof) {
    SC_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    SC_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    char line[320];
    snprintf(line, sizeof(line), "%s\t%.6f\t%.6e\t%.6e\n", r.name.c_str(), ms, r.flops, r.bytes);
    out += line;
  }
  *needed = out.size() + 1;
  if (buf && cap >= out.size() + 1) memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

int sparkcodec_launch_count(sparkcodec_handle* h, int64_t* count) {
  if (!h || !count) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  *count = h->launches;
  return 0;
}

int sparkcodec_fp32_terms(void) { return fp32_terms(); }

int sparkcodec_tile_width(int n_total, int cols_per_phase, int m_tiles, int num_sms) {
  if (n_total <= 0 || cols_per_phase <= 0 || n_total % cols_per_phase != 0 || m_tiles < 0 || num_sms <= 0) {
    set_error("tile_width: bad arguments");
    return SPARKCODEC_EINVAL;
  }
  int packed = 0;
  const int rc = choose_block_n(cols_per_phase, &packed);
  if (rc != 0) return rc;
  return launch_block_n(packed, n_total, cols_per_phase, m_tiles, num_sms);
}

int sparkcodec_pack_conv_f16f8(int kind, const float* w_host, const int64_t* wshape, int param, uint16_t* w_h16,
                               uint16_t* w_p8, size_t w_capacity) {
  if (!w_host || !wshape || !w_h16 || !w_p8) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  PackedGemm pk;
  if (kind == 0) pack_conv1d(w_host, (int)wshape[0], (int)wshape[1], (int)wshape[2], param, nullptr, nullptr, pk);
  else if (kind == 1) pack_conv_transpose1d(w_host, (int)wshape[0], (int)wshape[1], (int)wshape[2], param, nullptr, pk);
  else { set_error("kind must be 0 (Conv1d) or 1 (ConvTranspose1d)"); return SPARKCODEC_EINVAL; }
  if (pk.w_h16.empty()) { set_error("K = %d x %d is not a multiple of 32", pk.kt, pk.c_in); return SPARKCODEC_EINVAL; }
  if (pk.w_h16.size() > w_capacity) { set_error("w buffers too small (%zu needed)", pk.w_h16.size()); return SPARKCODEC_ENOMEM; }
  memcpy(w_h16, pk.w_h16.data(), pk.w_h16.size() * 2);
  memcpy(w_p8, pk.w_p8.data(), pk.w_p8.size() * 2);
  return 0;
}

int sparkcodec_pack_conv(int kind, const float* w_host, const int64_t* wshape, int param, uint16_t* w_hi,
                         uint16_t* w_lo, size_t w_capacity, int32_t* shifts, int32_t* ntaps, int32_t* kt,
                         int32_t* n_phase, int32_t* n_total) {
  if (!w_host || !wshape || !kt || !n_phase || !n_total) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  PackedGemm pk;
  if (kind == 0) pack_conv1d(w_host, (int)wshape[0], (int)wshape[1], (int)wshape[2], param, nullptr, nullptr, pk);
  else if (kind == 1) pack_conv_transpose1d(w_host, (int)wshape[0], (int)wshape[1], (int)wshape[2], param, nullptr, pk);
  else { set_error("kind must be 0 (Conv1d) or 1 (ConvTranspose1d)"); return SPARKCODEC_EINVAL; }
  *kt = pk.kt; *n_phase = pk.taps.n_phase; *n_total = pk.n_total;
  if (w_hi && w_lo) {
    if (pk.w_hi.size() > w_capacity) { set_error("w buffers too small (%zu needed)", pk.w_hi.size()); return SPARKCODEC_ENOMEM; }
    memcpy(w_hi, pk.w_hi.data(), pk.w_hi.size() * 2);
    memcpy(w_lo, pk.w_lo.data(), pk.w_lo.size() * 2);
  }
  if (ntaps) for (int r = 0; r < pk.taps.n_phase; ++r) ntaps[r] = pk.taps.ntaps[r];
  if (shifts)
    for (int r = 0; r < pk.taps.n_phase; ++r)
      for (int m = 0; m < pk.kt; ++m) shifts[r * pk.kt + m] = m < pk.taps.ntaps[r] ? pk.taps.shift[r][m] : INT32_MIN;
  return 0;
}

int sparkcodec_op_conv(int device, int kind, const float* w_host, const int64_t* wshape, const float* bias_host,
                       int param, int batch, int L, const float* x_dev, float* y_dev, const float* residual_dev,
                       int act, const float* alpha_host, int precision, int impl, void* stream) {
  if (!w_host || !wshape || !x_dev || !y_dev) { set_error("null argument"); return SPARKCODEC_EINVAL; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_ON_DEVICE(device);
  sparkcodec_handle tmp;   // only used as an allocation list
  tmp.device = device;
  cudaDeviceProp prop;
  SC_CUDA(cudaGetDeviceProperties(&prop, device));
  g_launch_counter = nullptr;
  PackedGemm pk;
  int c_in, c_out;
  if (kind == 0) {
    c_out = (int)wshape[0]; c_in = (int)wshape[1];
    pack_conv1d(w_host, c_out, c_in, (int)wshape[2], param, bias_host, nullptr, pk);
  } else if (kind == 1) {
    c_in = (int)wshape[0]; c_out = (int)wshape[1];
    pack_conv_transpose1d(w_host, c_in, c_out, (int)wshape[2], param, bias_host, pk);
  } else { set_error("kind must be 0 or 1"); return SPARKCODEC_EINVAL; }
  int rc = 0;
  auto cleanup = [&]() { cudaStreamSynchronize(st); for (void* p : tmp.allocs) cudaFree(p); };
  GemmWeights g;
  OpBuf a, o;
  a.fmt = o.fmt = op_fmt_for(precision);
  SnakeParams sp;
  void* p;
  const size_t n_in = (size_t)batch * L * c_in, n_out = (size_t)batch * L * pk.n_total;
  do {
    if ((rc = upload_gemm(&tmp, pk, &g))) break;
    if ((rc = dev_alloc(&tmp, n_in * 2, &p))) break; a.hi = (__nv_bfloat16*)p;
    if ((rc = dev_alloc(&tmp, n_in * 2, &p))) break; a.lo = (__nv_bfloat16*)p;
    if ((rc = launch_split(x_dev, a, n_in, st))) break;
    Epilogue e;
    e.residual = residual_dev;
    if (act == ACT_NONE) {
      e.out_f32 = y_dev;
    } else {
      if ((rc = dev_alloc(&tmp, n_out * 2, &p))) break; o.hi = (__nv_bfloat16*)p;
      if ((rc = dev_alloc(&tmp, n_out * 2, &p))) break; o.lo = (__nv_bfloat16*)p;
      e.act = act;
      e.out_op = o;
      if (act == ACT_SNAKE) {
        if (!alpha_host) { set_error("snake needs alpha"); rc = SPARKCODEC_EINVAL; break; }
        const int rep = pk.n_total / c_out;
        std::vector<float> al((size_t)pk.n_total), inv((size_t)pk.n_total);
        for (int r = 0; r < rep; ++r)
          for (int i = 0; i < c_out; ++i) {
            al[(size_t)r * c_out + i] = alpha_host[i];
            inv[(size_t)r * c_out + i] = 1.0f / (alpha_host[i] + 1e-9f);
          }
        if ((rc = upload(&tmp, al, &sp.alpha))) break;
        if ((rc = upload(&tmp, inv, &sp.inv))) break;
        e.alpha = sp.alpha; e.inv_alpha = sp.inv;
      }
    }
    if (!is_split(precision)) { a.lo = nullptr; }
    if (impl == SPARKCODEC_IMPL_SIMT) rc = launch_conv_gemm_simt(g, a, batch, L, e, precision, st);
    else rc = launch_conv_gemm_tc(g, a, batch, L, e, precision, prop.multiProcessorCount, st);
    if (rc) break;
    if (act != ACT_NONE) {
      OpBuf m = o;
      if (!is_split(precision)) m.lo = nullptr;
      if ((rc = launch_merge(m, y_dev, n_out, st))) break;
    }
    cudaError_t ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { set_error("op_conv: %s", cudaGetErrorString(ce)); rc = SPARKCODEC_ECUDA; }
  } while (0);
  cleanup();
  return rc;
}

}  // extern "C"
