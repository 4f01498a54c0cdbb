// kernels_beam.cu -- device-side best-first "beam" search bookkeeping
// (networks/EfficientSATRN.py:708-867, postprocessing/decoding.py:56-91).
//
// The reference keeps, per sample, a Python queue.PriorityQueue of
// (score, BeamSearchNode) tuples: score = -(logp / len) in fp64, ties broken by
// BeamSearchNode.__lt__ (len), anything beyond that by the heap's own structure.
// To reproduce its pop order exactly these kernels run the SAME binary-heap
// algorithm as CPython's heapq (_siftdown / _siftup) with the same comparison,
// one thread per sample, all samples of the batch advancing one expansion per
// round.  The expansion itself (one decoder step over the node's ancestor chain)
// is done by the batched step kernels with per-row positions / chains.
#include "common.cuh"
#include "kernels.h"

namespace frx {

namespace {

__device__ __forceinline__ bool item_less(const BeamP& p, int b, double sa, int na, double sb, int nb) {
  if (sa != sb) return sa < sb;                                  // tuple order: score first
  if (na == nb) return false;
  return p.nlen[(size_t)b * p.cap + na] < p.nlen[(size_t)b * p.cap + nb];  // BeamSearchNode.__lt__ (decoding.py:83-84)
}

// heapq.heappush
__device__ void heap_push(const BeamP& p, int b, double score, int node) {
  double* hs = p.hscore + (size_t)b * p.cap;
  int* hn = p.hnode + (size_t)b * p.cap;
  int pos = p.hsize[b]++;
  while (pos > 0) {  // _siftdown(heap, 0, pos)
    int parent = (pos - 1) >> 1;
    if (item_less(p, b, score, node, hs[parent], hn[parent])) {
      hs[pos] = hs[parent];
      hn[pos] = hn[parent];
      pos = parent;
    } else {
      break;
    }
  }
  hs[pos] = score;
  hn[pos] = node;
}

// heapq.heappop
__device__ int heap_pop(const BeamP& p, int b, double* score_out) {
  double* hs = p.hscore + (size_t)b * p.cap;
  int* hn = p.hnode + (size_t)b * p.cap;
  int n = --p.hsize[b];
  double ls = hs[n];
  int ln = hn[n];
  if (n == 0) { *score_out = ls; return ln; }
  double rs = hs[0];
  int rn = hn[0];
  // heap[0] = lastelt; _siftup(heap, 0)
  int pos = 0, child = 1;
  while (child < n) {
    int right = child + 1;
    if (right < n && !item_less(p, b, hs[child], hn[child], hs[right], hn[right])) child = right;
    hs[pos] = hs[child];
    hn[pos] = hn[child];
    pos = child;
    child = 2 * pos + 1;
  }
  // _siftdown(heap, 0, pos) with newitem = lastelt
  while (pos > 0) {
    int parent = (pos - 1) >> 1;
    if (item_less(p, b, ls, ln, hs[parent], hn[parent])) {
      hs[pos] = hs[parent];
      hn[pos] = hn[parent];
      pos = parent;
    } else {
      break;
    }
  }
  hs[pos] = ls;
  hn[pos] = ln;
  *score_out = rs;
  return rn;
}

}  // namespace

// Root node (<SOS>, logp 0, len 1) with score -(0/1) (:736-750).
__global__ void beam_init_kernel(const BeamP p) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  size_t o = (size_t)b * p.cap;
  p.nprev[o] = -1; p.ntok[o] = p.sos; p.nlen[o] = 1; p.nlogp[o] = 0.0; p.nkv[o] = -1;
  p.ncount[b] = 1;
  p.hsize[b] = 0;
  heap_push(p, b, -(0.0 / 1.0), 0);
  p.num_steps[b] = 0; p.done[b] = 0; p.endnode[b] = -1; p.nexp[b] = 0; p.active[b] = 0;
  p.cur_tok[b] = p.sos; p.pos[b] = 0; p.slot[b] = p.T - 1; p.cur_node[b] = 0;
}

// One round of the `while True` loop up to the decoder step (:753-771).
__global__ void beam_select_kernel(const BeamP p) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  p.active[b] = 0;
  p.slot[b] = p.T - 1;  // inactive rows scatter their K/V into a scratch slot
  p.pos[b] = 0;
  if (p.done[b]) return;
  size_t o = (size_t)b * p.cap;
  double score;
  if (p.num_steps[b] >= (p.max_seq - 1) * p.bw) {  // budget exhausted (:754): best remaining node (:834-835)
    p.endnode[b] = heap_pop(p, b, &score);
    p.done[b] = 1;
    return;
  }
  int n = heap_pop(p, b, &score);
  if (p.ntok[o + n] == p.eos && p.nprev[o + n] != -1) {  // first popped EOS ends the search (:764-767)
    p.endnode[b] = n;
    p.done[b] = 1;
    return;
  }
  p.cur_node[b] = n;
  p.cur_tok[b] = p.ntok[o + n];
  const int pos = p.nlen[o + n] - 1;  // position for the 1-D PE (:761)
  p.pos[b] = pos;
  int* chain = p.chain + (size_t)b * p.T;
  for (int a = p.nprev[o + n]; a != -1; a = p.nprev[o + a]) chain[p.nlen[o + a] - 1] = p.nkv[o + a];
  const int e = p.nexp[b]++;
  p.nkv[o + n] = e;  // this expansion's layer outputs (as K/V rows) live in cache row e
  p.slot[b] = e;
  p.active[b] = 1;
  atomicAdd(p.n_active, 1);
}

// log_softmax + topk(beam_width) + child nodes (:806-831); one warp per sample.
__global__ void __launch_bounds__(256) beam_push_kernel(const BeamP p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= p.B || !p.active[b]) return;
  const float* lg = p.logits + (size_t)b * p.V;
  float x[8];  // V <= 256
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int v = i * 32 + lane;
    x[i] = v < p.V ? lg[v] : -INFINITY;
    mx = fmaxf(mx, x[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i * 32 + lane < p.V) sum += expf(x[i] - mx);
  sum = warp_sum(sum);
  const float lse = logf(sum);
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (x[i] - mx) - lse;  // log_softmax
  size_t o = (size_t)b * p.cap;
  const int n = p.cur_node[b];
  for (int k = 0; k < p.bw; ++k) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int v = i * 32 + lane;
      if (v < p.V && (x[i] > best)) { best = x[i]; bi = v; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, s);
      int oi = __shfl_xor_sync(0xffffffffu, bi, s);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (bi >= 0 && bi < p.V && (bi & 31) == lane) x[bi >> 5] = -INFINITY;  // remove from further rounds
    if (lane == 0 && bi >= 0 && bi < p.V) {
      const int c = p.ncount[b]++;
      p.nprev[o + c] = n;
      p.ntok[o + c] = bi;
      p.nlen[o + c] = p.nlen[o + n] + 1;
      const double logp = p.nlogp[o + n] + (double)best;  // n.logp + log_p (python float, :821)
      p.nlogp[o + c] = logp;
      p.nkv[o + c] = -1;
      heap_push(p, b, -(logp / (double)p.nlen[o + c]), c);  // score = -node.eval() (decoding.py:80)
    }
    __syncwarp();
  }
  if (lane == 0) p.num_steps[b] += p.bw;  // :831
}

// Back-trace (:838-849) and PAD / truncate to max_sequence (:857-865).
__global__ void beam_finish_kernel(const BeamP p) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  size_t o = (size_t)b * p.cap;
  long long* out = p.out + (size_t)b * p.max_seq;
  int end = p.endnode[b];
  int len = end >= 0 ? p.nlen[o + end] : 0;
  for (int i = 0; i < p.max_seq; ++i) out[i] = p.pad;
  for (int a = end; a != -1 && a >= 0; a = p.nprev[o + a]) {
    int at = p.nlen[o + a] - 1;
    if (at < p.max_seq) out[at] = p.ntok[o + a];
  }
  (void)len;
}

void launch_beam_init(const BeamP& p, cudaStream_t st) { beam_init_kernel<<<(p.B + 127) / 128, 128, 0, st>>>(p); }
void launch_beam_select(const BeamP& p, cudaStream_t st) { beam_select_kernel<<<(p.B + 127) / 128, 128, 0, st>>>(p); }
void launch_beam_push(const BeamP& p, cudaStream_t st) { beam_push_kernel<<<(p.B + 7) / 8, 256, 0, st>>>(p); }
void launch_beam_finish(const BeamP& p, cudaStream_t st) { beam_finish_kernel<<<(p.B + 127) / 128, 128, 0, st>>>(p); }

__global__ void fill_i32_kernel(int* p, int v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
void launch_fill_i32(int* p, int v, int n, cudaStream_t st) { fill_i32_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, v, n); }

}  // namespace frx
