// ensemble.cu -- the ensemble decoding loop of utils/ensemble_utils.py:71-118 (make_decoder_values) on the device.
//
// Every model of the ensemble is one frx handle (EfficientSATRN_decoder: step_forward / reset_status,
// networks/EfficientSATRN.py:932-952).  Per step the reference calls model.step_forward(src_m, target) for every model,
// averages F.softmax of their logits (:95-105), optionally passes the average through DecodingManager.sift (:107-108),
// takes the arg-max as the next target (:110) and appends the distribution (:112).  Here the encoder memories, the
// per-model logits, the averaged distribution and the next target all stay in HBM: one averaging kernel per step,
// no host synchronisation inside the loop.
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/frx.h"
#include "kernels.h"
#include "runtime.h"

using namespace frx;

namespace {

constexpr int ENS_MAX = 8;
struct EnsPtrs { const float* p[ENS_MAX]; };

// out[b][:] = mean_m softmax(logits_m[b][:]); one warp per row.  Without a manager also the arg-max (first maximum, like
// torch.argmax) -> target[b] and tokens[b * ld_tok].
__global__ void __launch_bounds__(256) ens_average_kernel(EnsPtrs lg, int n_models, int B, int V, float* __restrict__ out, long long ld_out,
                                                          long long* __restrict__ target, long long* __restrict__ tokens, long long ld_tok,
                                                          int pick) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  float acc[8];   // V <= 256
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int m = 0; m < n_models; ++m) {
    const float* lp = lg.p[m] + (long long)b * V;
    float v[8], mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; v[i] = c < V ? lp[c] : -INFINITY; mx = fmaxf(mx, v[i]); }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = lane + 32 * i < V ? expf(v[i] - mx) : 0.f; s += v[i]; }
    s = warp_sum(s);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += v[i] / s;       // one_step_out += F.softmax(_out) (:101-103)
  }
  float best = -INFINITY;
  int bi = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c >= V) continue;
    const float pv = acc[i] / (float)n_models;            // :105
    out[(long long)b * ld_out + c] = pv;
    if (pv > best) { best = pv; bi = c; }
  }
  if (!pick) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) {
    if (bi < 0 || bi >= V) bi = 0;
    target[b] = bi;
    if (tokens) tokens[(long long)b * ld_tok] = bi;
  }
}

__global__ void ens_fill_target_kernel(long long* target, int B, long long v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) target[i] = v;
}
__global__ void ens_gather_target_kernel(const int* __restrict__ cur_tok, long long* __restrict__ target, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) target[i] = cur_tok[i];
}

int efail(frx_handle* h, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return 1;
}

}  // namespace

extern "C" int frx_ensemble_decode(frx_handle* const* handles, int32_t n_models, const float* const* memories, int32_t B, int32_t steps,
                                   float* probs, int64_t* tokens, int32_t use_manager, void* stream) {
  if (!handles || n_models < 1 || !handles[0]) return 1;
  frx_handle* h0 = handles[0];
  if (n_models > ENS_MAX) return efail(h0, "ensemble_decode: at most %d models", ENS_MAX);
  if (!probs) return efail(h0, "ensemble_decode: probs is required (the averaged distributions are the result)");
  const int V = h0->cfg.num_classes;
  if (V > 256) return efail(h0, "ensemble_decode: more than 256 classes not supported");
  for (int m = 0; m < n_models; ++m) {
    frx_handle* h = handles[m];
    if (!h || !memories || !memories[m]) return efail(h0, "ensemble_decode: model %d: null handle or memory", m);
    if (h->cfg.num_classes != V || h->cfg.device != h0->cfg.device) return efail(h0, "ensemble_decode: model %d: vocabulary / device differ from model 0", m);
    if (steps < 1 || steps > h->cfg.max_steps || B < 1 || B > h->cfg.max_batch)
      return efail(h0, "ensemble_decode: model %d: batch %d / steps %d outside its limits (%d, %d)", m, B, steps, h->cfg.max_batch, h->cfg.max_steps);
  }
  if (use_manager && !h0->sift_flags) return efail(h0, "ensemble_decode: no decoding rules set on model 0 (frx_set_decoding_rules)");
  cudaStream_t st = (cudaStream_t)stream;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != h0->cfg.device) cudaSetDevice(h0->cfg.device);
  int rc = 0;
  long long* target = nullptr;
  float* lbuf = nullptr;
  do {
    if (cudaMallocAsync((void**)&target, (size_t)B * 8, st) != cudaSuccess ||
        cudaMallocAsync((void**)&lbuf, (size_t)n_models * B * V * 4, st) != cudaSuccess) { rc = efail(h0, "ensemble_decode: out of device memory"); break; }
    EnsPtrs lp{};
    for (int m = 0; m < n_models; ++m) {
      lp.p[m] = lbuf + (size_t)m * B * V;
      if (frx_decode_begin(handles[m], memories[m], B, st)) { if (handles[m] != h0) h0->err = "model " + std::to_string(m) + ": " + handles[m]->err; rc = 1; break; }
    }
    if (rc) break;
    ens_fill_target_kernel<<<(B + 255) / 256, 256, 0, st>>>(target, B, h0->cfg.sos_id);      // align <SOS> (:81-82)
    const SiftIds sids{h0->sift_ids[0], h0->sift_ids[1], h0->sift_ids[2], h0->sift_ids[3], h0->sift_ids[4], h0->sift_ids[5]};
    if (use_manager) launch_sift_state_init(h0->sift_state, B, sids.sos, st);                 // manager.reset (:85-86)
    for (int t = 0; t < steps && !rc; ++t) {
      for (int m = 0; m < n_models; ++m)
        if (frx_decode_step(handles[m], (const int64_t*)target, const_cast<float*>(lp.p[m]), st)) { if (handles[m] != h0) h0->err = "model " + std::to_string(m) + ": " + handles[m]->err; rc = 1; break; }
      if (rc) break;
      float* out = probs + (size_t)t * V;
      const long long ld = (long long)steps * V;
      ens_average_kernel<<<(B + 7) / 8, 256, 0, st>>>(lp, n_models, B, V, out, ld, target, tokens ? (long long*)tokens + t : nullptr, steps,
                                                      use_manager ? 0 : 1);
      h0->launches++;
      if (use_manager) {
        // manager.sift(one_step_out) (:107-108): the manager soft-maxes its input AGAIN (postprocessing.py:216) -- the
        // averaged probabilities are treated as logits, exactly as the reference does
        launch_dec_sift_embed(out, ld, V, tokens ? (long long*)tokens + t : nullptr, steps, h0->sift_state, h0->sift_flags, h0->sift_limit, sids,
                              h0->cur_tok, nullptr, nullptr, 0.f, nullptr, B, h0->cfg.dec_hidden, st);
        ens_gather_target_kernel<<<(B + 255) / 256, 256, 0, st>>>(h0->cur_tok, target, B);
        h0->launches += 2;
      }
    }
    const cudaError_t e = cudaGetLastError();
    if (!rc && e != cudaSuccess) rc = efail(h0, "ensemble_decode: kernel launch failed: %s", cudaGetErrorString(e));
  } while (0);
  if (target) cudaFreeAsync(target, st);
  if (lbuf) cudaFreeAsync(lbuf, st);
  for (int m = 0; m < n_models; ++m) { handles[m]->step_idx = 0; handles[m]->step_batch = 0; }   // model.reset_status() (:117-118)
  if (prev != h0->cfg.device && prev >= 0) cudaSetDevice(prev);
  return rc;
}
