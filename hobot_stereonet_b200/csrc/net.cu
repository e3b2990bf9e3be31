// Weight loading, scratch arena and the static op plan of the StereoNet forward pass.
// Topology: SURVEY.md §2.3 (reconstructed from the reference's .hbm tensor table); the free
// choices are documented in DESIGN.md §2 and restated identically in oracle/stereonet_ref.py.
#include "net.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "kernels.cuh"
#include "layers.h"

namespace snb {

static const int LAYER_BLOCKS[4] = {3, 16, 3, 3};
static const int LAYER_CH[4] = {32, 64, 128, 128};
static const int REF_DIL[6] = {1, 2, 4, 8, 1, 1};
// zero border (pixels) of the padded C8 layout: >= the largest dilation of any tcgen05 consumer
static const int PAD_BACKBONE = 2, PAD_REFINE = 16;   // the fused residual block reads a 2*dilation halo

// ---- weight blob ("SNB2WGT1", oracle/weights.py documents the layout) -----------------------------
// The blob named by `model_file` is untrusted input: every offset and size is range-checked without overflow, every
// tensor's byte count must equal 4 * prod(dims), and the table must hold exactly the convolutions of the topology
// build_plan assumes (layers.h), with those shapes.  Nothing of the context is touched: the result goes to `out`.
int parse_blob(const void* blob, size_t bytes, std::map<std::string, HostTensor>* out, int* blob_K, char* err, size_t errlen) {
  const uint8_t* p = static_cast<const uint8_t*>(blob);
  if (!p || bytes < 24 || memcmp(p, "SNB2WGT1", 8) != 0) {
    snprintf(err, errlen, "model_file is not a SNB2WGT1 weight blob");
    return SNB_ERR_MODEL;
  }
  uint32_t ver, K, n;
  memcpy(&ver, p + 8, 4); memcpy(&K, p + 12, 4); memcpy(&n, p + 16, 4);
  if (ver != 1 || K < 1 || K > 5 || n > 4096 || bytes < 24 + (size_t)n * 104) {
    snprintf(err, errlen, "weight blob: bad version / K or truncated table");
    return SNB_ERR_MODEL;
  }
  const size_t base = (24 + (size_t)n * 104 + 63) / 64 * 64;
  if (base > bytes) { snprintf(err, errlen, "weight blob: truncated before the data section"); return SNB_ERR_MODEL; }
  const size_t avail = bytes - base;
  std::map<std::string, HostTensor> wts;
  for (uint32_t i = 0; i < n; ++i) {
    const uint8_t* e = p + 24 + (size_t)i * 104;
    char name[65]; memcpy(name, e, 64); name[64] = 0;
    uint32_t ndim, dims[5]; uint64_t off, nb;
    memcpy(&ndim, e + 64, 4); memcpy(dims, e + 68, 20); memcpy(&off, e + 88, 8); memcpy(&nb, e + 96, 8);
    if (off > avail || nb > avail - off) {
      snprintf(err, errlen, "weight blob: tensor %s out of range", name);
      return SNB_ERR_MODEL;
    }
    uint64_t cnt = 1;
    bool ok = ndim == 1 || ndim == 4 || ndim == 5;
    for (uint32_t k = 0; ok && k < ndim; ++k) {
      if (dims[k] == 0 || dims[k] > 65536) ok = false;
      else cnt *= dims[k];
      if (cnt > ((uint64_t)1 << 32)) ok = false;
    }
    if (!ok || nb != cnt * 4) {
      snprintf(err, errlen, "weight blob: tensor %s has a bad shape or byte count", name);
      return SNB_ERR_MODEL;
    }
    HostTensor t;
    for (uint32_t k = 0; k < ndim; ++k) t.shape.push_back((int)dims[k]);
    t.data.resize(cnt);
    memcpy(t.data.data(), p + base + off, nb);
    wts[name] = std::move(t);
  }
  // the topology build_plan assumes: every convolution present, with these shapes
  for (const Spec& sp : conv_specs((int)K)) {
    auto wi = wts.find(sp.name + ".weight"), bi = wts.find(sp.name + ".bias");
    if (wi == wts.end() || bi == wts.end()) {
      snprintf(err, errlen, "weight blob lacks %s", sp.name.c_str());
      return SNB_ERR_MODEL;
    }
    std::vector<int> want = {sp.cout, sp.cin};
    if (sp.kd) want.push_back(sp.kd);
    want.push_back(sp.ks); want.push_back(sp.ks);
    if (wi->second.shape != want || bi->second.shape != std::vector<int>{sp.cout}) {
      snprintf(err, errlen, "weight blob: %s does not have the shape of the network (Cout %d, Cin %d, k %d%s)", sp.name.c_str(),
               sp.cout, sp.cin, sp.ks, sp.kd ? ", 3-D" : "");
      return SNB_ERR_MODEL;
    }
  }
  *out = std::move(wts);
  *blob_K = (int)K;
  return SNB_OK;
}

static int pack_conv(snb_ctx* c, const std::string& name) {
  auto wi = c->wts.find(name + ".weight"), bi = c->wts.find(name + ".bias");
  if (wi == c->wts.end() || bi == c->wts.end()) {
    snprintf(c->err, sizeof(c->err), "weight blob lacks %s", name.c_str());
    return SNB_ERR_MODEL;
  }
  const HostTensor& W = wi->second;
  ConvW cw;
  cw.cout = W.shape[0]; cw.cin = W.shape[1];
  const bool is3d = W.shape.size() == 5;
  cw.kz = is3d ? W.shape[2] : 1;
  cw.ks = W.shape.back();
  cw.wlog2 = weight_scale_log2(W.data.data(), W.data.size());
  const int ntap = cw.ks * cw.ks, cbin = (cw.cin + 7) / 8;
  std::vector<float> packed;
  if (cw.cout == 1) {
    packed.assign((size_t)cbin * cw.kz * 9 * 8, 0.f);
    for (int ci = 0; ci < cw.cin; ++ci)
      for (int kz = 0; kz < cw.kz; ++kz)
        for (int t = 0; t < ntap; ++t)
          packed[(((size_t)(ci / 8) * cw.kz + kz) * 9 + t) * 8 + ci % 8] =
              W.data[(((size_t)ci) * cw.kz + kz) * ntap + t];
    cw.b0 = bi->second.data[0];
  } else {
    const int CO = (cw.cout % 32 == 0) ? 32 : 16;
    const int ncc = cw.cout / CO;
    packed.assign((size_t)ncc * cbin * cw.kz * ntap * 8 * CO, 0.f);
    for (int co = 0; co < cw.cout; ++co)
      for (int ci = 0; ci < cw.cin; ++ci)
        for (int kz = 0; kz < cw.kz; ++kz)
          for (int t = 0; t < ntap; ++t) {
            const size_t dst = (((((size_t)(co / CO) * cbin + ci / 8) * cw.kz + kz) * ntap + t) * 8 + ci % 8) * CO + co % CO;
            packed[dst] = W.data[((((size_t)co * cw.cin + ci) * cw.kz + kz) * ntap) + t];
          }
  }
  if (cudaMalloc(&cw.w, packed.size() * 4) != cudaSuccess || cudaMalloc(&cw.b, cw.cout * 4) != cudaSuccess) {
    snprintf(c->err, sizeof(c->err), "cudaMalloc weights failed");
    return SNB_ERR_NOMEM;
  }
  c->wallocs.push_back(cw.w); c->wallocs.push_back(cw.b);
  cudaMemcpy(cw.w, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(cw.b, bi->second.data.data(), cw.cout * 4, cudaMemcpyHostToDevice);
  c->convs[name] = cw;
  return SNB_OK;
}

int upload_weights(snb_ctx* c) {
  for (void* p : c->wallocs) cudaFree(p);
  c->wallocs.clear(); c->convs.clear();
  if (c->blob_K != c->K) {
    snprintf(c->err, sizeof(c->err), "weight blob was generated for K=%d, config asks K=%d", c->blob_K, c->K);
    return SNB_ERR_MODEL;
  }
  for (const Spec& sp : conv_specs(c->K)) {
    int r = pack_conv(c, sp.name);
    if (r != SNB_OK) return r;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return SNB_ERR_CUDA;
  return SNB_OK;
}

// ---- scratch arena: stream-ordered reuse of activation buffers -----------------------------------
// A block is only ever reused by a tensor of identical geometry (`key`): the zero borders of the padded
// C8 layout are written once, here, and every kernel writes interior pixels only.
void* Arena::get(size_t bytes, uint64_t key) {
  bytes = (bytes + 255) / 256 * 256;
  if (reuse) {
    for (auto& b : blks)
      if (b.free && b.key == key && b.bytes == bytes) { b.free = false; return b.p; }
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  cudaMemset(p, 0, bytes);
  blks.push_back({p, bytes, false, key});
  total += bytes;
  return p;
}
void Arena::put(void* p) {
  if (!reuse) return;
  for (auto& b : blks) if (b.p == p) b.free = true;
}
void Arena::release_all() {
  for (auto& b : blks) cudaFree(b.p);
  blks.clear(); total = 0;
}

// ---- plan builder --------------------------------------------------------------------------------
struct Builder {
  snb_ctx* c;
  int maxB;
  bool fail = false;

  Tens alloc(int nmul, int ch, int d, int h, int w, int pad) {
    Tens t; t.n = nmul * maxB; t.c = ch; t.cb = (ch + 7) / 8; t.d = d; t.h = h; t.w = w; t.planes = c->planes; t.pad = pad;
    const uint64_t key = ((((uint64_t)t.n * 64 + t.cb) * 1024 + t.d) * 8192 + t.h) * 8192 + t.w + ((uint64_t)pad << 58);
    t.p = c->arena.get(t.bytes(), key);
    if (!t.p) fail = true;
    return t;
  }
  Plane palloc(int d, int h, int w) {
    Plane p; p.n = maxB; p.d = d; p.h = h; p.w = w;
    p.p = static_cast<float*>(c->arena.get(p.bytes(), ~(uint64_t)0));
    if (!p.p) fail = true;
    return p;
  }
  void free(const Tens& t) { if (!t.pcb) c->arena.put(t.p); }    // sub-views do not own memory
  void free(const Plane& p) { c->arena.put(p.p); }
  void tap(const std::string& name, const Tens& t, int nmul) { Stage s; s.t = t; s.nmul = nmul; c->stages[name] = s; }
  void tap(const std::string& name, const Plane& p) { Stage s; s.is_plane = true; s.p = p; c->stages[name] = s; }

  // tcgen05 path: stride-1 3x3 / 3x3x3 convolutions with Cin % 16 == 0 and Cout % 32 == 0
  bool tc_eligible(const ConvW& cw, int stride) const {
    return c->planes == 2 && !(c->cfg.flags & SNB_FLAG_NO_TENSOR) && stride == 1 && cw.ks == 3 && (cw.cin % 16 == 0 || (cw.cin <= 8 && cw.kz == 1)) && cw.cout % 32 == 0;
  }

  // device copy of the tcgen05 packing of one convolution's weights (made once, shared by every user)
  const __half* tc_weights(const std::string& name, ConvW& cw) {
    const int NT = cw.cin <= 8 ? 8 : 32;   // packing id: 32 = 16-channel K chunks, 8 = "cin8" tap-pair packing
    if (!cw.w_tc.count(NT)) {
      std::vector<__half> packed;
      tc_pack_weights(c->wts[name + ".weight"].data.data(), cw.cout, cw.cin, cw.kz, NT, cw.wlog2, packed);
      __half* dw = nullptr;
      if (cudaMalloc(&dw, packed.size() * sizeof(__half)) != cudaSuccess) { fail = true; return nullptr; }
      cudaMemcpy(dw, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice);
      c->wallocs.push_back(dw);
      cw.w_tc[NT] = dw;
    }
    return cw.w_tc[NT];
  }

  // device copy of the k_conv_stream packing (NCO = 32: C8 output slices, 16: single output channel)
  const __half* stream_weights(const std::string& name, ConvW& cw, int nco) {
    const int key = 100 + nco;
    if (!cw.w_tc.count(key)) {
      std::vector<__half> packed;
      cs_pack_weights(c->wts[name + ".weight"].data.data(), cw.cout, cw.cin, cw.kz, cw.ks, nco, cw.wlog2, packed);
      __half* dw = nullptr;
      if (cudaMalloc(&dw, packed.size() * sizeof(__half)) != cudaSuccess) { fail = true; return nullptr; }
      cudaMemcpy(dw, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice);
      c->wallocs.push_back(dw);
      cw.w_tc[key] = dw;
    }
    return cw.w_tc[key];
  }
  // k_conv_stream packing of input channels [c0, c0 + nc) of a convolution (a K split across launches)
  const __half* stream_weights_part(const std::string& name, ConvW& cw, int c0, int nc) {
    const int key = 1000 + c0;
    if (!cw.w_tc.count(key)) {
      const std::vector<float>& W = c->wts[name + ".weight"].data;
      const size_t taps = (size_t)cw.kz * cw.ks * cw.ks;
      std::vector<float> part((size_t)cw.cout * nc * taps);
      for (int co = 0; co < cw.cout; ++co)
        std::copy(W.begin() + ((size_t)co * cw.cin + c0) * taps, W.begin() + ((size_t)co * cw.cin + c0 + nc) * taps, part.begin() + (size_t)co * nc * taps);
      std::vector<__half> packed;
      cs_pack_weights(part.data(), cw.cout, nc, cw.kz, cw.ks, 32, cw.wlog2, packed);
      __half* dw = nullptr;
      if (cudaMalloc(&dw, packed.size() * sizeof(__half)) != cudaSuccess) { fail = true; return nullptr; }
      cudaMemcpy(dw, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice);
      c->wallocs.push_back(dw);
      cw.w_tc[key] = dw;
    }
    return cw.w_tc[key];
  }
  const float* zero_bias() {
    if (!zbias) {
      if (cudaMalloc(&zbias, 32 * sizeof(float)) != cudaSuccess) { fail = true; return nullptr; }
      cudaMemset(zbias, 0, 32 * sizeof(float));
      c->wallocs.push_back(zbias);
    }
    return zbias;
  }
  float* zbias = nullptr;
  bool stream_ok(const ConvW& cw, const Tens& in, int stride, int dil, CsPlan* plan) const {
    if (c->planes != 2 || (c->cfg.flags & (SNB_FLAG_NO_TENSOR | SNB_FLAG_NO_STREAM)) || stride > 2 || (stride == 2 && cw.cout == 1)) return false;
    // input channels as stored (blocks of 8, zero beyond cin) must be whole 16-channel chunks
    if (in.cb * 8 < cw.cin || (in.cb * 8) % 16) return false;
    if (conv_stream_plan(plan, in, in.cb * 8, cw.cout, cw.ks == 1 ? 1 : dil, cw.kz, c->num_sms) != cudaSuccess) return false;
    conv_stream_set_taps(plan, cw.ks);
    // measured on config 2 (profiles/r01_final_opprof.txt): with the weights of one slice resident the streaming kernel
    // beats the staged k_conv_tc tiles wherever the slice fits (layer3/4 27 vs 30 us, head.filter.1-4 43 vs 50 us,
    // firstconv.1 26 vs 39, conv_out 41 vs 52, conv3d_alone 39 vs 64); head.filter.0 (221 KB) and lastconv.0 (295 KB) do not fit
    static const int max_kb = getenv("SNB_STREAM_MAX_KB") ? atoi(getenv("SNB_STREAM_MAX_KB")) : 160;
    return cw.cout == 1 || plan->p.w_bytes <= (uint32_t)max_kb * 1024;
  }

  // out = ReLU(conv_b(ReLU(conv_a(in))) + res) with 32 channels, as ONE launch when the tcgen05 path allows it
  bool block_fusable(const std::string& prefix, const Tens& in, int dil) const {
    auto ia = c->convs.find(prefix + ".conv_a"), ib = c->convs.find(prefix + ".conv_b");
    if (ia == c->convs.end() || ib == c->convs.end()) return false;
    if (ia->second.cin != 32 || ia->second.cout != 32 || ib->second.cin != 32 || ib->second.cout != 32 || ia->second.kz != 1) return false;
    return c->planes == 2 && !(c->cfg.flags & (SNB_FLAG_NO_TENSOR | SNB_FLAG_NO_FUSE)) && in.c == 32 && in.d == 1 && in.pad >= 2 * dil;
  }
  Tens resblock(const std::string& prefix, const Tens& in, int nmul, int dil, const Tens& res) {
    auto ia = c->convs.find(prefix + ".conv_a"), ib = c->convs.find(prefix + ".conv_b");
    if (ia == c->convs.end() || ib == c->convs.end()) { fail = true; snprintf(c->err, sizeof(c->err), "no weights for %s", prefix.c_str()); return Tens(); }
    Tens out = alloc(nmul, 32, 1, in.h, in.w, in.pad);
    RbPlan plan;
    cudaError_t e = resblock_tc_plan(&plan, in, out, res, dil, c->num_sms);
    if (e != cudaSuccess) { fail = true; snprintf(c->err, sizeof(c->err), "resblock_tc_plan(%s): %s", prefix.c_str(), cudaGetErrorString(e)); return out; }
    const __half* wa = tc_weights(prefix + ".conv_a", ia->second);
    const __half* wb = tc_weights(prefix + ".conv_b", ib->second);
    const float* ba = ia->second.b; const float* bb = ib->second.b;
    Op op; op.name = prefix + " [tc-block]";
    const double px = (double)nmul * in.h * in.w;
    op.flops = 2.0 * 2.0 * px * 32 * 32 * 9;
    op.bytes = 4.0 * px * 32 * (&res == &in || res.p == in.p ? 2 : 3);
    const int la = ia->second.wlog2, lb = ib->second.wlog2;
    op.fn = [plan, nmul, wa, wb, la, lb, ba, bb](int B, const IoPtrs&, cudaStream_t st) { return launch_resblock_tc(plan, nmul * B, wa, wb, la, lb, ba, bb, st); };
    c->n_tc_convs += 2;
    c->ops.push_back(op);
    return out;
  }

  // conv_a (3x3, ReLU) and the 1x1 shortcut convolution of a BasicBlock read the same input: ONE streaming launch serves both,
  // the shortcut as extra output slices (k_conv_stream second head).  false: not applicable, the caller emits two launches.
  bool conv_with_shortcut(const std::string& name_a, const std::string& name_s, const Tens& in, int nmul, int stride, int dil, Tens* out_a,
                          Tens* out_s) {
    auto ia = c->convs.find(name_a), is = c->convs.find(name_s);
    if (ia == c->convs.end() || is == c->convs.end() || (c->cfg.flags & SNB_FLAG_NO_FUSE)) return false;
    ConvW& ca = ia->second; ConvW& cs = is->second;
    if (ca.ks != 3 || cs.ks != 1 || ca.kz != 1 || cs.kz != 1 || ca.cin != cs.cin || ca.cout % 32 || cs.cout % 32) return false;
    CsPlan splan;
    if (!stream_ok(ca, in, stride, dil, &splan)) return false;
    const int ho = (in.h + stride - 1) / stride, wo = (in.w + stride - 1) / stride;
    const Tens oa = alloc(nmul, ca.cout, in.d, ho, wo, in.pad), os = alloc(nmul, cs.cout, in.d, ho, wo, in.pad);
    const __half* wa = stream_weights(name_a, ca, 32);
    const __half* ws = stream_weights(name_s, cs, 32);
    if (!wa || !ws) return false;
    const float* ba = ca.b; const float* bs = cs.b;
    const int la = ca.wlog2, ls = cs.wlog2, couts = cs.cout;
    Op op; op.name = name_a + " + " + name_s.substr(name_s.rfind('.') + 1) + " [tc-stream x2]";
    const double px = (double)nmul * in.d * ho * wo;
    op.flops = 2.0 * px * (ca.cout * ca.cin * 9.0 + cs.cout * cs.cin);
    op.bytes = 4.0 * (nmul * (double)in.d * in.h * in.w * in.cb * 8 + px * (oa.cb + os.cb) * 8);
    op.fn = [splan, nmul, wa, la, ba, oa, ws, ls, bs, os, couts, stride](int B, const IoPtrs&, cudaStream_t st) {
      CsHead2 h2{ws, ls, bs, &os, couts, 1, 0};
      return launch_conv_stream(splan, nmul * B, wa, la, ba, &oa, nullptr, nullptr, nullptr, 0, 1, stride, st, 2, nullptr, 0, 0, 0.f, &h2);
    };
    c->n_tc_convs += 2;
    c->ops.push_back(op);
    *out_a = oa; *out_s = os;
    return true;
  }

  Tens conv(const std::string& name, const Tens& in, int nmul, int stride, int dil, bool relu, const Tens* res,
            const Tens* dst = nullptr) {
    auto it = c->convs.find(name);
    if (it == c->convs.end()) { fail = true; snprintf(c->err, sizeof(c->err), "no weights for %s", name.c_str()); return Tens(); }
    ConvW& cw = it->second;
    const int ho = (in.h + stride - 1) / stride, wo = (in.w + stride - 1) / stride;
    Tens out = dst ? *dst : alloc(nmul, cw.cout, in.d, ho, wo, in.pad);
    Op op; op.name = name;
    const double px = (double)nmul * in.d * ho * wo;
    op.flops = 2.0 * px * cw.cout * cw.cin * cw.ks * cw.ks * cw.kz;
    op.bytes = 4.0 * (nmul * (double)in.d * in.h * in.w * in.cb * 8 + px * out.cb * 8 * (res ? 2 : 1));
    if (c->direct_io && cw.cout == 32 && cw.ks == 3 && cw.kz == 1 && cw.cin == 3 && stride == 2 && !res && dil == 1) {
      // firstconv.0 straight from the s8 model input (k_conv_first_s8): `in` is only a geometry carrier here, it has no memory
      ConvFirstS8Params fp{};
      fp.out = view(out); fp.H = c->H; fp.W = c->W; fp.Ho = ho; fp.Wo = wo; fp.relu = relu ? 1 : 0;
      conv_first_s8_pack(c->wts[name + ".weight"].data.data(), c->wts[name + ".bias"].data.data(), &fp);
      op.fn = [fp](int B, const IoPtrs& io, cudaStream_t st) { return launch_conv_first_s8(fp, io, B, st); };
      op.name += " [s8]";
      op.io_bytes = (int)sizeof(ConvFirstS8Params);
      op.bytes = 1.0 * nmul * 3.0 * c->H * c->W + 4.0 * px * out.cb * 8;       // three s8 planes per view in, C8 split out
      c->ops.push_back(op);
      return out;
    }
    if (c->planes == 2 && !(c->cfg.flags & SNB_FLAG_NO_HBMCONV) && cw.cout == 32 && cw.ks == 3 && cw.kz == 1 && in.d == 1 && in.pad >= 1 && !res &&
        dil == 1 && cw.cin == 3 && stride == 2) {
      // the 3-channel image convolution: 13 FLOP per byte moved, CUDA cores with the weights in the constant bank (k_conv_hbm.cu);
      // the 4-channel refinement conv_in measured 38.6 us this way against 40.0 us on k_conv_tc (FMA bound) and stays there
      ConvFirstParams fp{};
      fp.in = view(in); fp.out = view(out); fp.Ho = ho; fp.Wo = wo; fp.relu = relu ? 1 : 0; fp.stride = stride;
      conv_first_pack(c->wts[name + ".weight"].data.data(), c->wts[name + ".bias"].data.data(), cw.cin, &fp);
      op.fn = [fp, nmul](int B, const IoPtrs&, cudaStream_t st) { return launch_conv_first(fp, nmul * B, st); };
      op.name += " [hbm]";
      op.bytes = (cw.cin == 3 ? 2.0 : 4.0) * nmul * (double)in.h * in.w * 8 + 4.0 * px * out.cb * 8;   // one channel block in, C8 split out
      c->ops.push_back(op);
      return out;
    }
    if (c->planes == 2 && !(c->cfg.flags & (SNB_FLAG_NO_TENSOR | SNB_FLAG_NO_HBMCONV)) && cw.ks == 1 && cw.kz == 1 && in.d == 1 && stride == 1 && !res &&
        (cw.cout == 16 || cw.cout == 32) && in.cb * 8 >= cw.cin) {
      // a 1x1 convolution on its own (layer1.0's shortcut, lastconv.1): far too little work for a tensor-pipe launch (k_conv1x1)
      std::vector<float> packed;
      conv1x1_pack(c->wts[name + ".weight"].data.data(), cw.cout, cw.cin, in.cb, packed);
      float* dw = nullptr;
      if (cudaMalloc(&dw, packed.size() * sizeof(float)) != cudaSuccess) { fail = true; return out; }
      cudaMemcpy(dw, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice);
      c->wallocs.push_back(dw);
      Conv1x1Params qp{};
      qp.in = view(in); qp.out = view(out); qp.wgt = dw; qp.bias = cw.b; qp.h = in.h; qp.w = in.w; qp.cbin = in.cb; qp.relu = relu ? 1 : 0;
      const int cout = cw.cout;
      op.fn = [qp, cout, nmul](int B, const IoPtrs&, cudaStream_t st) { return launch_conv1x1(qp, cout, nmul * B, st); };
      op.name += " [1x1]";
      c->ops.push_back(op);
      return out;
    }
    CsPlan splan;
    static const bool ksplit_ok = !getenv("SNB_STREAM_KSPLIT") || atoi(getenv("SNB_STREAM_KSPLIT"));
    if (ksplit_ok && cw.kz == 3 && cw.ks == 3 && cw.cin == 64 && cw.cout == 32 && in.cb == 8 && stride == 1 && !res && !stream_ok(cw, in, stride, dil, &splan)) {
      // head.filter.0 (Conv3d 64 -> 32): its 221 KB weight slice does not fit beside a row ring, but each HALF of the input channels
      // does - two streaming launches, the second adding the first's output, run at the rate of head.filter.1-4 (428 TFLOP/s at
      // D = 192 against 299 for the staged tiles of k_conv_tc).  The partial sum passes through the split-fp16 storage (22 bits).
      ConvW half = cw; half.cin = 32;
      const Tens in0 = sub_view(in, 0, 4), in1 = sub_view(in, 4, 4);
      CsPlan p0, p1;
      if (stream_ok(half, in0, 1, dil, &p0) && stream_ok(half, in1, 1, dil, &p1)) {
        const __half* w0 = stream_weights_part(name, cw, 0, 32);
        const __half* w1 = stream_weights_part(name, cw, 32, 32);
        const float* zb = zero_bias();
        if (!w0 || !w1 || !zb) return out;
        const Tens mid = alloc(nmul, cw.cout, in.d, ho, wo, in.pad);
        const float* bias = cw.b;
        const int rl = relu ? 1 : 0, wl = cw.wlog2;
        Op o0 = op, o1 = op;
        o0.flops = o1.flops = op.flops / 2;
        o0.bytes = 4.0 * (nmul * (double)in.d * in.h * in.w * 32 + px * out.cb * 8);
        o1.bytes = 4.0 * (nmul * (double)in.d * in.h * in.w * 32 + 2 * px * out.cb * 8);
        o0.name += " [tc-stream, input channels 0-31]";
        o1.name += " [tc-stream, input channels 32-63 + the first half]";
        o0.fn = [p0, nmul, w0, wl, bias, mid](int B, const IoPtrs&, cudaStream_t st) {
          return launch_conv_stream(p0, nmul * B, w0, wl, bias, &mid, nullptr, nullptr, nullptr, 0, 0, 1, st);
        };
        o1.fn = [p1, nmul, w1, wl, zb, out, mid, rl](int B, const IoPtrs&, cudaStream_t st) {
          return launch_conv_stream(p1, nmul * B, w1, wl, zb, &out, &mid, nullptr, nullptr, 0, rl, 1, st);
        };
        ++c->n_tc_convs;
        c->ops.push_back(o0);
        c->ops.push_back(o1);
        free(mid);
        return out;
      }
    }
    if (stream_ok(cw, in, stride, dil, &splan)) {
      const __half* dw = stream_weights(name, cw, 32);
      if (!dw) return out;
      const float* bias = cw.b;
      const bool has_res = res != nullptr;
      const Tens rt = res ? *res : Tens();
      const int rl = relu ? 1 : 0;
      const int wl = cw.wlog2;
      op.fn = [splan, nmul, dw, wl, bias, out, has_res, rt, rl, stride](int B, const IoPtrs&, cudaStream_t st) {
        return launch_conv_stream(splan, nmul * B, dw, wl, bias, &out, has_res ? &rt : nullptr, nullptr, nullptr, 0, rl, stride, st);
      };
      op.name += " [tc-stream]";
      ++c->n_tc_convs;
      c->ops.push_back(op);
      return out;
    }
    if (tc_eligible(cw, stride) && in.pad >= dil) {
      TcConvPlan plan;
      cudaError_t e = tc_conv_plan(&plan, in, out, cw.cin, cw.cout, dil, cw.kz, c->num_sms);
      if (e != cudaSuccess) { fail = true; snprintf(c->err, sizeof(c->err), "tc_conv_plan(%s): %s", name.c_str(), cudaGetErrorString(e)); return out; }
      const __half* dw = tc_weights(name, cw);
      if (!dw) return out;
      const float* bias = cw.b;
      const bool has_res = res != nullptr;
      const Tens rt = res ? *res : Tens();
      const int sms = c->num_sms, rl = relu ? 1 : 0;
      const int wl = cw.wlog2;
      op.fn = [plan, nmul, dw, wl, bias, has_res, rt, rl, sms](int B, const IoPtrs&, cudaStream_t st) {
        return launch_conv_tc(plan, nmul * B, dw, wl, bias, has_res ? &rt : nullptr, rl, sms, st);
      };
      op.name += " [tc]";
      ++c->n_tc_convs;
      c->ops.push_back(op);
      return out;
    }
    ConvParams p{};
    p.in = view(in); p.out = view(out); p.w = cw.w; p.bias = cw.b;
    if (res) p.res = view(*res);
    p.CBin = std::min(in.cb, (cw.cin + 7) / 8); p.Din = in.d; p.Hin = in.h; p.Win = in.w;
    p.CBout = out.cb; p.Dout = out.d; p.Hout = ho; p.Wout = wo;
    p.ks = cw.ks; p.kz = cw.kz; p.stride = stride; p.dil = dil; p.relu = relu ? 1 : 0;
    p.half = c->planes == 2;
    const int cout = cw.cout;
    op.fn = [p, nmul, cout](int B, const IoPtrs&, cudaStream_t st) mutable { ConvParams q = p; q.N = nmul * B; return launch_conv_direct(q, cout, st); };
    ++c->n_direct_convs;
    c->ops.push_back(op);
    return out;
  }

  // res_up: the residual is the x2 bilinear upsample of this half-resolution plane (direct_io refinement: no 4-channel input tensor
  // exists to take it from); emit_q: this is the last conv_out and also writes the s32 model output
  Plane conv_to1(const std::string& name, const Tens& in, int dil, bool relu, const Tens* res_c8, const Plane* res_up = nullptr,
                 bool emit_q = false) {
    auto it = c->convs.find(name);
    if (it == c->convs.end()) { fail = true; snprintf(c->err, sizeof(c->err), "no weights for %s", name.c_str()); return Plane(); }
    ConvW& cw = it->second;
    Plane out = palloc(in.d, in.h, in.w);
    CsPlan splan;
    if (cw.cin % 16 == 0 && cw.ks == 3 && stream_ok(cw, in, 1, dil, &splan)) {
      const __half* dw = stream_weights(name, cw, 16);
      if (!dw) return out;
      const float* bias = cw.b;
      const bool has_res = res_c8 != nullptr;
      const Tens rt = res_c8 ? *res_c8 : Tens();
      const int rl = relu ? 1 : 0;
      float* op_ = out.p;
      Op op; op.name = name + (cw.kz == 3 && cw.cin == 32 && !res_c8 && !res_up && !relu && !emit_q ? " [tc-cost3d]" : " [tc-stream]");
      const int wl = cw.wlog2;
      const float* rup = res_up ? res_up->p : nullptr;
      const int qH = c->H, qW = c->W;
      const float qmul = c->qmul;
      // Conv3d 32 -> 1 (conv3d_alone): the read-once kernel (k_cost3d.cu).  Measured against the streaming kernel at config 2:
      // 22 vs 30 us at one pair per pass, 120 vs 215 us at eight; D = 192, four pairs: 333 vs 623 us - and its shorter
      // accumulator chains leave less truncation bias in the cost tensor (SNB_COST3D_MIN_ROWS=<huge> restores the streaming kernel)
      Cost3dPlan c3{};
      const __half* dw3 = nullptr;
      static const int c3_min_rows = getenv("SNB_COST3D_MIN_ROWS") ? atoi(getenv("SNB_COST3D_MIN_ROWS")) : 0;
      if (cw.kz == 3 && cw.cin == 32 && !res_c8 && !res_up && !relu && !emit_q && cost3d_plan(&c3, in, c->num_sms) == cudaSuccess) {
        const int key3 = 300;
        if (!cw.w_tc.count(key3)) {
          std::vector<__half> packed;
          cost3d_pack_weights(c->wts[name + ".weight"].data.data(), cw.cin, cw.wlog2, packed);
          __half* d3 = nullptr;
          if (cudaMalloc(&d3, packed.size() * sizeof(__half)) != cudaSuccess) { fail = true; return out; }
          cudaMemcpy(d3, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice);
          c->wallocs.push_back(d3);
          cw.w_tc[key3] = d3;
        }
        dw3 = cw.w_tc[key3];
      }
      const float b0 = cw.b0;
      const long rows_per_pair = (long)in.d * in.h;
      op.fn = [splan, dw, wl, bias, op_, has_res, rt, rl, rup, emit_q, qH, qW, qmul, c3, dw3, b0, rows_per_pair](int B, const IoPtrs& io, cudaStream_t st) {
        if (dw3 && rows_per_pair * B >= c3_min_rows) return launch_cost3d(c3, B, dw3, wl, b0, op_, st);
        return launch_conv_stream(splan, B, dw, wl, bias, nullptr, has_res ? &rt : nullptr, op_, rup, 1, rl, 1, st, 3, emit_q ? &io : nullptr, qH, qW,
                                  qmul);
      };
      if (emit_q) { op.io_bytes = (int)sizeof(CsParams); op.name += " [+s32 out]"; }
      const double px = (double)in.d * in.h * in.w;
      op.flops = 2.0 * px * cw.cin * 9 * cw.kz;
      op.bytes = 4.0 * (px * in.cb * 8 + px * (res_c8 ? 2 : 1));
      ++c->n_tc_convs;
      c->ops.push_back(op);
      return out;
    }
    if (res_up || emit_q) { fail = true; snprintf(c->err, sizeof(c->err), "%s: the fused refinement path needs the streaming kernel", name.c_str()); return out; }
    ConvTo1Params p{};
    p.in = view(in); p.out = out.p; p.w = cw.w; p.bias = cw.b0;
    p.res_c8 = res_c8 ? 1 : 0;
    if (res_c8) p.res = view(*res_c8);
    p.CBin = in.cb; p.D = in.d; p.H = in.h; p.W = in.w; p.kz = cw.kz; p.dil = dil; p.relu = relu ? 1 : 0;
    p.half = c->planes == 2;
    Op op; op.name = name;
    op.fn = [p](int B, const IoPtrs&, cudaStream_t st) { ConvTo1Params q = p; q.N = B; return launch_conv_to1(q, st); };
    const double px = (double)in.d * in.h * in.w;
    op.flops = 2.0 * px * cw.cin * 9 * cw.kz;
    op.bytes = 4.0 * (px * in.cb * 8 + px * (res_c8 ? 2 : 1));
    c->ops.push_back(op);
    return out;
  }
};

int build_plan(snb_ctx* c) {
  c->ops.clear(); c->stages.clear();
  c->arena.release_all();
  c->arena.reuse = !(c->cfg.flags & SNB_FLAG_KEEP_STAGES);
  c->planes = c->cfg.precision == SNB_PREC_TC_F16X2 ? 2 : 1;
  c->n_tc_convs = c->n_direct_convs = 0;
  // Tensor-core path: the kernels of the pass read the s8 model input themselves (firstconv.0, the refinement heads) and the
  // last conv_out writes the s32 model output: no image tensor, no pre-process / post-process launch.  SNB_FLAG_NO_HEADFUSE
  // (and the other diagnostic paths) keep the older pipeline: pre kernel -> C8 image -> ... -> soft-argmin, refine_in, conv_in.
  c->direct_io = c->planes == 2 && !(c->cfg.flags & (SNB_FLAG_NO_TENSOR | SNB_FLAG_NO_STREAM | SNB_FLAG_NO_HBMCONV | SNB_FLAG_NO_HEADFUSE));
  Builder b{c, c->maxB};
  const int K = c->K, D = c->D, Hp = c->Hp, Wp = c->Wp, h = c->h, w = c->w;

  Tens img;
  if (c->direct_io) {
    c->img = Tens();                     // geometry only: firstconv.0 reads the s8 tensor
    img.n = 2 * c->maxB; img.c = 3; img.cb = 1; img.h = Hp; img.w = Wp; img.planes = 2; img.pad = PAD_BACKBONE;
  } else {
    // input image, C8 [2B][1][Hp][Wp][8]; the pre-process op is issued by the caller (s8 or NV12 source)
    // 16 stored channels (3 real) on the tensor-core path: firstconv.0 consumes whole 16-channel K chunks
    c->img = b.alloc(2, c->planes == 2 ? 16 : 3, 1, Hp, Wp, PAD_BACKBONE);
    c->img.c = 3;
    img = c->img;
    b.tap("img", img, 2);
  }

  // ---- M1 siamese backbone (left and right batched as n = 2B) ----
  Tens x = b.conv("backbone.firstconv.0", img, 2, 2, 1, true, nullptr);
  Tens y = b.conv("backbone.firstconv.1", x, 2, 1, 1, true, nullptr); b.free(x);
  x = b.conv("backbone.firstconv.2", y, 2, 2, 1, true, nullptr); b.free(y);
  b.tap("firstconv", x, 2);
  const int strides[4] = {1, K >= 3 ? 2 : 1, K >= 4 ? 2 : 1, 1};
  // gwc feature = cat(layer3, layer4) along channels, [2B][32][h][w][8]: the last block of each layer writes its half
  Tens gwc = b.alloc(2, 256, 1, h, w, PAD_BACKBONE);
  const Tens gwc_l3 = sub_view(gwc, 0, 16), gwc_l4 = sub_view(gwc, 16, 16);
  Tens l3, l4;
  for (int li = 1; li <= 4; ++li) {
    for (int bi = 0; bi < LAYER_BLOCKS[li - 1]; ++bi) {
      const std::string p = "backbone.layer" + std::to_string(li) + "." + std::to_string(bi);
      const int s = bi == 0 ? strides[li - 1] : 1, dil = li == 4 ? 2 : 1;
      Tens sc = x;
      const bool ds = bi == 0 && li <= 3;
      const bool fused_block = s == 1 && b.block_fusable(p, x, dil);
      Tens a;
      bool have_a = false;
      if (ds) {
        if (!fused_block && b.conv_with_shortcut(p + ".conv_a", p + ".downsample", x, 2, s, dil, &a, &sc)) have_a = true;
        else sc = b.conv(p + ".downsample", x, 2, s, 1, false, nullptr);
      }
      Tens o;
      if (fused_block) {
        o = b.resblock(p, x, 2, dil, sc);
      } else {
        const bool last = bi == LAYER_BLOCKS[li - 1] - 1;
        const Tens* dst = (li == 3 && last) ? &gwc_l3 : (li == 4 && last) ? &gwc_l4 : nullptr;
        if (!have_a) a = b.conv(p + ".conv_a", x, 2, s, dil, true, nullptr);
        o = b.conv(p + ".conv_b", a, 2, 1, dil, true, &sc, dst);
        b.free(a);
      }
      if (ds) b.free(sc);
      b.free(x);
      x = o;
    }
    b.tap("layer" + std::to_string(li), x, 2);
    if (li == 3) l3 = x;
    if (li == 4) l4 = x;
  }
  (void)LAYER_CH;
  b.tap("gwc", gwc, 2);
  Tens lc = b.conv("backbone.lastconv.0", gwc, 2, 1, 1, true, nullptr);
  Tens cat = b.conv("backbone.lastconv.1", lc, 2, 1, 1, false, nullptr); b.free(lc);
  b.tap("cat", cat, 2);

  // ---- M2 cost volume [B][8][D][h][w][8] ----
  Tens vol = b.alloc(1, 64, D, h, w, PAD_BACKBONE);
  {
    Op op; op.name = "costvol";
    op.fn = [=](int B, const IoPtrs&, cudaStream_t st) { return launch_costvol(gwc, cat, vol, B, D, st); };
    op.flops = 2.0 * 256 * D * h * w;
    op.bytes = 4.0 * (2.0 * (256 + 16) * h * w + 64.0 * D * h * w);
    c->ops.push_back(op);
  }
  b.free(gwc); b.free(cat);
  b.tap("volume", vol, 1);

  // ---- M3 3-D aggregation ----
  Tens v = vol;
  for (int i = 0; i < 5; ++i) {
    Tens o = b.conv("head.filter." + std::to_string(i), v, 1, 1, 1, true, nullptr);
    b.free(v); v = o;
    b.tap("filter" + std::to_string(i), v, 1);
  }
  Plane cost = b.conv_to1("head.conv3d_alone", v, 1, false, nullptr); b.free(v);
  b.tap("cost", cost);

  // ---- M4 soft-argmin + M5 edge-aware refinement x K ----
  Plane disp = b.palloc(1, h, w);
  if (!c->direct_io) {
    Op op; op.name = "softargmin";
    op.fn = [=](int B, const IoPtrs&, cudaStream_t st) { Plane cc = cost; cc.n = B; return launch_softargmin(cc, disp, st); };
    op.bytes = 4.0 * (D + 1.0) * h * w;
    c->ops.push_back(op);
    b.free(cost);
  }
  b.tap("disp0", disp);

  for (int s = 0; s < K; ++s) {
    const std::string p = "head.refine." + std::to_string(s);
    const int hs = disp.h * 2, ws = disp.w * 2;
    Tens rin, f;
    if (c->direct_io) {
      // one launch: (stage 0: soft-argmin over D ->) x2 bilinear(disp) || left image from the s8 input -> conv_in + ReLU (k_refine_head)
      auto wi = c->wts.find(p + ".conv_in.weight"), bi = c->wts.find(p + ".conv_in.bias");
      if (wi == c->wts.end() || bi == c->wts.end()) { b.fail = true; snprintf(c->err, sizeof(c->err), "no weights for %s.conv_in", p.c_str()); break; }
      f = b.alloc(1, 32, 1, hs, ws, PAD_REFINE);
      RefineHeadParams rp{};
      rp.src = s == 0 ? cost.p : disp.p; rp.disp0 = disp.p; rp.out = view(f);
      rp.D = D; rp.h = disp.h; rp.w = disp.w; rp.H = c->H; rp.W = c->W; rp.f = Hp / hs; rp.stage0 = s == 0 ? 1 : 0; rp.invD = 1.0f / D;
      refine_head_pack(wi->second.data.data(), bi->second.data.data(), &rp);
      Op op; op.name = p + (s == 0 ? ".head [softargmin+up+conv_in]" : ".head [up+conv_in]");
      op.fn = [rp](int B, const IoPtrs& io, cudaStream_t st) { return launch_refine_head(rp, io, B, st); };
      op.io_bytes = (int)sizeof(RefineHeadParams);
      op.flops = 2.0 * hs * ws * 32 * 4 * 9;
      // algorithmic HBM bytes: cost (stage 0) or coarse disparity in, the left image at this resolution (3 s8 planes), split-fp16 feature out
      op.bytes = (s == 0 ? 4.0 * (D + 1.0) * disp.h * disp.w : 4.0 * disp.h * disp.w) + 3.0 * std::min((double)c->H * c->W, (double)hs * ws * (rp.f == 1 ? 1 : 4)) +
                 4.0 * 32.0 * hs * ws;
      c->ops.push_back(op);
      if (s == 0) b.free(cost);
    } else {
      rin = b.alloc(1, 4, 1, hs, ws, PAD_REFINE);
      Op op; op.name = p + ".in";
      Plane dsrc = disp;
      op.fn = [=](int B, const IoPtrs&, cudaStream_t st) { return launch_refine_in(dsrc, img, rin, B, st); };
      op.bytes = 4.0 * (disp.h * disp.w + 3.0 * hs * ws + 8.0 * hs * ws);
      c->ops.push_back(op);
      f = b.conv(p + ".conv_in", rin, 1, 1, 1, true, nullptr);
    }
    for (int bi = 0; bi < 6; ++bi) {
      const std::string q = p + ".blocks." + std::to_string(bi);
      Tens o;
      if (b.block_fusable(q, f, REF_DIL[bi])) {
        o = b.resblock(q, f, 1, REF_DIL[bi], f);
      } else {
        Tens a = b.conv(q + ".conv_a", f, 1, 1, REF_DIL[bi], true, nullptr);
        o = b.conv(q + ".conv_b", a, 1, 1, REF_DIL[bi], true, &f);
        b.free(a);
      }
      b.free(f); f = o;
    }
    b.tap("refine" + std::to_string(s) + ".feat", f, 1);
    Plane nd = c->direct_io ? b.conv_to1(p + ".conv_out", f, 1, true, nullptr, &disp, s == K - 1) : b.conv_to1(p + ".conv_out", f, 1, true, &rin);
    b.free(f); if (!c->direct_io) b.free(rin);
    b.free(disp);
    disp = nd;
    b.tap("disp" + std::to_string(s + 1), disp);
  }
  c->disp_final = disp;
  if (b.fail) {
    if (!c->err[0]) snprintf(c->err, sizeof(c->err), "device allocation failed while building the plan");
    return SNB_ERR_NOMEM;
  }
  return SNB_OK;
}

// The pass for batch B on stream st, reading / writing the buffers in `io`.  With use_graph the pass is captured once per
// (B, kind of entry) and replayed; a later call with other buffers re-points the few kernels that touch them
// (cudaGraphExecKernelNodeSetParams on the nodes recorded at capture: CPU work only, nothing is added on the stream).
int run_plan(snb_ctx* c, int B, const IoPtrs& io, cudaStream_t st, bool use_graph) {
  if (use_graph) {
    const int key = 2 * B + (io.frames ? 1 : 0);
    auto it = c->graphs.find(key);
    if (it == c->graphs.end()) {
      GraphEntry ge;
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return SNB_ERR_CUDA;
      cudaError_t e = cudaSuccess;
      std::vector<std::pair<cudaGraphNode_t, int>> io_nodes;
      auto note_io_node = [&](int bytes) {     // the node just captured = the stream's only dependency now
        cudaStreamCaptureStatus status; const cudaGraphNode_t* deps = nullptr; size_t ndeps = 0;
        cudaError_t ce = cudaStreamGetCaptureInfo(st, &status, nullptr, nullptr, &deps, &ndeps);
        if (ce != cudaSuccess || status != cudaStreamCaptureStatusActive || ndeps != 1) return ce == cudaSuccess ? cudaErrorUnknown : ce;
        io_nodes.push_back({deps[0], bytes});
        return cudaSuccess;
      };
      if (io.frames && c->direct_io) {         // camera-frame entry: P1-P3 on the GPU is the first kernel of the pass
        e = launch_pre_nv12(io.frames, Tens(), const_cast<int8_t*>(io.s8), B, c->H, c->W, (c->cfg.flags & SNB_FLAG_CORRECT_CHROMA) ? 1 : 0, st);
        if (e == cudaSuccess) e = note_io_node((int)sizeof(PreNv12Params));
      }
      for (auto& op : c->ops) {
        if (e != cudaSuccess) break;
        e = op.fn(B, io, st);
        if (e == cudaSuccess && op.io_bytes > 0) e = note_io_node(op.io_bytes);
      }
      cudaError_t e2 = cudaStreamEndCapture(st, &g);
      if (e != cudaSuccess || e2 != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "graph capture failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        if (g) cudaGraphDestroy(g);
        return SNB_ERR_CUDA;
      }
      for (auto& pr : io_nodes) {
        std::unique_ptr<GraphEntry::IoNode> n(new GraphEntry::IoNode());
        n->node = pr.first;
        if (cudaGraphKernelNodeGetParams(pr.first, &n->kp) != cudaSuccess || !n->kp.kernelParams) { cudaGraphDestroy(g); return SNB_ERR_CUDA; }
        n->args.assign(static_cast<const char*>(n->kp.kernelParams[0]), static_cast<const char*>(n->kp.kernelParams[0]) + pr.second);
        n->argv[0] = n->args.data();
        n->kp.kernelParams = n->argv; n->kp.extra = nullptr;
        ge.nodes.push_back(std::move(n));
      }
      if (cudaGraphInstantiate(&ge.exec, g, 0) != cudaSuccess) { cudaGraphDestroy(g); return SNB_ERR_CUDA; }
      ge.graph = g;
      ge.io = io;
      it = c->graphs.emplace(key, std::move(ge)).first;
    }
    GraphEntry& ge = it->second;
    if (ge.io.s8 != io.s8 || ge.io.q != io.q || ge.io.frames != io.frames) {
      for (auto& n : ge.nodes) {
        memcpy(n->args.data(), &io, sizeof(IoPtrs));       // IoPtrs is the first member of every io kernel's parameter struct
        if (cudaGraphExecKernelNodeSetParams(ge.exec, n->node, &n->kp) != cudaSuccess) {
          snprintf(c->err, sizeof(c->err), "cudaGraphExecKernelNodeSetParams: %s", cudaGetErrorString(cudaGetLastError()));
          return SNB_ERR_CUDA;
        }
      }
      ge.io = io;
    }
    if (cudaGraphLaunch(ge.exec, st) != cudaSuccess) {
      snprintf(c->err, sizeof(c->err), "cudaGraphLaunch: %s", cudaGetErrorString(cudaGetLastError()));
      return SNB_ERR_CUDA;
    }
    return SNB_OK;
  }
  if (io.frames && c->direct_io) {
    cudaError_t e = launch_pre_nv12(io.frames, Tens(), const_cast<int8_t*>(io.s8), B, c->H, c->W, (c->cfg.flags & SNB_FLAG_CORRECT_CHROMA) ? 1 : 0, st);
    if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "pre_nv12: %s", cudaGetErrorString(e)); return SNB_ERR_CUDA; }
  }
  for (auto& op : c->ops) {
    cudaError_t e = op.fn(B, io, st);
    if (e != cudaSuccess) {
      snprintf(c->err, sizeof(c->err), "%s: %s", op.name.c_str(), cudaGetErrorString(e));
      return SNB_ERR_CUDA;
    }
  }
  return SNB_OK;
}

void free_ctx(snb_ctx* c) {
  for (auto& g : c->graphs) { cudaGraphExecDestroy(g.second.exec); if (g.second.graph) cudaGraphDestroy(g.second.graph); }
  c->graphs.clear();
  c->arena.release_all();
  for (void* p : c->wallocs) cudaFree(p);
  c->wallocs.clear();
  for (auto& sl : c->slots) {
    if (sl.d_in) cudaFree(sl.d_in);
    if (sl.d_frames) cudaFree(sl.d_frames);
    if (sl.d_out) cudaFree(sl.d_out);
    if (sl.e_in) cudaEventDestroy(sl.e_in);
    if (sl.e_done) cudaEventDestroy(sl.e_done);
    if (sl.e_out) cudaEventDestroy(sl.e_out);
  }
  c->slots.clear();
  if (c->st_in) cudaStreamDestroy(c->st_in);
  if (c->st_out) cudaStreamDestroy(c->st_out);
  if (c->d_in) cudaFree(c->d_in);
  if (c->d_out) cudaFree(c->d_out);
  if (c->d_frames) cudaFree(c->d_frames);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  if (c->ev_last) cudaEventDestroy(c->ev_last);
  if (c->stream) cudaStreamDestroy(c->stream);
}

}  // namespace snb
