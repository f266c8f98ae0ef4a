// TokenLabelCrossEntropy: dense token-label CE + class-token CE, forward AND gradient in one pass.
//
// Replaces loss/cross_entropy.py:136-156 (+ SoftTargetCrossEntropy :30-36):
//   L = (w_cls/B) sum_b CE(x_cls[b], t_cls[b]) + (w_dense/(B N)) sum_{b,n} CE(x_aux[b,n], target[b,:,2+n])
//   CE(x,t) = -sum_c t_c log_softmax(x)_c ;  dL/dx = w * (softmax(x) * sum_c t_c - t)            
//   t_cls[b] = lam * target[b,:,1] + (1-lam) * target[B-1-b,:,1]  when lam = 1 - box_area/N < 1   (:149-151)
// The target is CLASS-major [B, C, 2+N] while the logits are token-major [B, N, C]; a CTA stages
// TT tokens x C classes of both in shared memory (one HBM read of each), so the transpose happens
// on chip and each logit / target element is read once and each gradient element written once.
#include "common.cuh"

namespace {

constexpr int TT = 16;        // tokens per CTA
constexpr int TS = TT + 1;    // padded token stride of the staged target tile (bank-conflict free transposed reads)
constexpr int NTHREADS = 256;

struct TlceParams {
  const void* x_cls;   // [B, C]
  const void* x_aux;   // [B, N, C]
  const float* target; // 3-D: [B, C, 2+N]   2-D: [B, C]
  void* d_cls;
  void* d_aux;
  float* partial;      // [gridDim.x] per-CTA loss partials (already weighted)
  int B, N, C;
  long long t_sb, t_sc, t_ss;  // target strides (batch, class, slot); 2-D target: (C, 1, 0)
  int slot_cls, slot_aux0;     // 3-D: 1, 2   2-D: 0, 0
  float lam;                   // cls-target mix factor (>= 1 -> no mixing)
  const int* box_dev;          // optional device (bbx1,bby1,bbx2,bby2): overrides lam (CUDA-graph path)
  float w_cls, w_dense;        // already divided by B and B*N
  int tiles_per_img;
  int cta_offset;              // added to blockIdx.x (the fast path launches only the class-token CTAs of tlce_kernel)
  // single-launch fast path (ticket != nullptr): CTAs [0, B) class tokens, [B, B + n_aux) dense tiles; the LAST CTA
  // to finish (ticket counter) sums all partials in the fixed order of tlce_reduce_kernel, writes the loss, re-arms the ticket
  int* ticket;
  float* loss;
  int n_aux;
};

template <typename T>
__global__ void __launch_bounds__(NTHREADS) tlce_kernel(TlceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, N = p.N;
  __shared__ float s_loss[NTHREADS / 32];
  float my_loss = 0.f;
  const int n_aux_ctas = p.B * p.tiles_per_img;
  const int bid = (int)blockIdx.x + p.cta_offset;

  if (bid < n_aux_ctas) {
    // ---------------- dense part: TT tokens of one image ----------------
    const int b = bid / p.tiles_per_img;
    const int n0 = (bid % p.tiles_per_img) * TT;
    const int nt = min(TT, N - n0);
    float* st = reinterpret_cast<float*>(smem_raw);                       // [C][TS]
    T* sx = reinterpret_cast<T*>(smem_raw + (size_t)C * TS * sizeof(float));  // [TT][C]
    const T* xg = reinterpret_cast<const T*>(p.x_aux) + ((size_t)b * N + n0) * C;
    T* dg = reinterpret_cast<T*>(p.d_aux) + ((size_t)b * N + n0) * C;
    // logits tile: nt*C contiguous elements
    const size_t tot = (size_t)nt * C;
    constexpr int VN = Vec16<T>::N;
    if ((C % VN) == 0) {
      for (size_t e = (size_t)tid * VN; e < tot; e += (size_t)NTHREADS * VN) {
        Vec16<T> v;
        v.load(xg + e);
        v.store(sx + e);
      }
    } else {
      for (size_t e = tid; e < tot; e += NTHREADS) sx[e] = xg[e];
    }
    // target tile: for each class, nt consecutive slots
    const float* tg = p.target + (size_t)b * p.t_sb + (size_t)(p.slot_aux0 + n0) * p.t_ss;
    for (int e = tid; e < C * TT; e += NTHREADS) {
      const int c = e / TT, n = e % TT;
      st[c * TS + n] = (n < nt) ? tg[(size_t)c * p.t_sc + (size_t)n * p.t_ss] : 0.f;
    }
    __syncthreads();
    for (int n = warp; n < nt; n += NTHREADS / 32) {
      const T* xr = sx + (size_t)n * C;
      float m = -INFINITY;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, to_f(xr[c]));
      m = warp_max(m);
      float se = 0.f, sum_t = 0.f, sum_tx = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float x = to_f(xr[c]), t = st[c * TS + n];
        se += expf(x - m);
        sum_t += t;
        sum_tx = fmaf(t, x, sum_tx);
      }
      se = warp_sum(se);
      sum_t = warp_sum(sum_t);
      sum_tx = warp_sum(sum_tx);
      const float lse = m + logf(se);
      if (lane == 0) my_loss += p.w_dense * (lse * sum_t - sum_tx);
      T* dr = dg + (size_t)n * C;
      for (int c = lane; c < C; c += 32) {
        const float x = to_f(xr[c]), t = st[c * TS + n];
        dr[c] = from_f<T>(p.w_dense * (expf(x - lse) * sum_t - t));
      }
    }
  } else {
    // ---------------- class-token part: one image per warp ----------------
    const int b0 = (bid - n_aux_ctas) * (NTHREADS / 32);
    const int b = b0 + warp;
    if (b < p.B) {
      const T* xr = reinterpret_cast<const T*>(p.x_cls) + (size_t)b * C;
      T* dr = reinterpret_cast<T*>(p.d_cls) + (size_t)b * C;
      const float* t0 = p.target + (size_t)b * p.t_sb + (size_t)p.slot_cls * p.t_ss;
      const float* t1 = p.target + (size_t)(p.B - 1 - b) * p.t_sb + (size_t)p.slot_cls * p.t_ss;
      float lam = p.lam;
      if (p.box_dev != nullptr)
        lam = 1.f - (float)((p.box_dev[2] - p.box_dev[0]) * (p.box_dev[3] - p.box_dev[1])) / (float)p.N;
      const bool mix = lam < 1.f;
      float m = -INFINITY;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, to_f(xr[c]));
      m = warp_max(m);
      float se = 0.f, sum_t = 0.f, sum_tx = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float x = to_f(xr[c]);
        float t = t0[(size_t)c * p.t_sc];
        if (mix) t = lam * t + (1.f - lam) * t1[(size_t)c * p.t_sc];
        se += expf(x - m);
        sum_t += t;
        sum_tx = fmaf(t, x, sum_tx);
      }
      se = warp_sum(se);
      sum_t = warp_sum(sum_t);
      sum_tx = warp_sum(sum_tx);
      const float lse = m + logf(se);
      if (lane == 0) my_loss += p.w_cls * (lse * sum_t - sum_tx);
      for (int c = lane; c < C; c += 32) {
        const float x = to_f(xr[c]);
        float t = t0[(size_t)c * p.t_sc];
        if (mix) t = lam * t + (1.f - lam) * t1[(size_t)c * p.t_sc];
        dr[c] = from_f<T>(p.w_cls * (expf(x - lse) * sum_t - t));
      }
    }
  }
  if (lane == 0) s_loss[warp] = my_loss;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < NTHREADS / 32; ++i) s += s_loss[i];
    p.partial[bid] = s;
  }
}

// Class-token part of the fast path: one CTA per image, thread = classes tid, tid + 256, ...  The class-level labels sit
// at stride (2 + N) floats in the class-major target, i.e. every load is its own sector: all of a thread's loads (own image
// and, when mixing, the flipped image) are issued before the first use and kept in registers for the gradient pass -- the
// warp-per-image loop this replaces walked 2 x 32 dependent strided loads per lane (32 us for 128 images).
constexpr int CLS_MAXJ = 4;            // classes per thread (C <= 1024)
template <typename T>
__device__ __forceinline__ void tlce_cls_body(const TlceParams& p, int b, int partial_index) {
  __shared__ float red[3][NTHREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C;
  const T* xr = reinterpret_cast<const T*>(p.x_cls) + (size_t)b * C;
  T* dr = reinterpret_cast<T*>(p.d_cls) + (size_t)b * C;
  const float* t0 = p.target + (size_t)b * p.t_sb + (size_t)p.slot_cls * p.t_ss;
  const float* t1 = p.target + (size_t)(p.B - 1 - b) * p.t_sb + (size_t)p.slot_cls * p.t_ss;
  float lam = p.lam;
  if (p.box_dev != nullptr) lam = 1.f - (float)((p.box_dev[2] - p.box_dev[0]) * (p.box_dev[3] - p.box_dev[1])) / (float)p.N;
  const bool mix = lam < 1.f;
  float x[CLS_MAXJ], t[CLS_MAXJ], u[CLS_MAXJ];
#pragma unroll
  for (int j = 0; j < CLS_MAXJ; ++j) {
    const int c = tid + j * NTHREADS;
    const bool ok = c < C;
    x[j] = ok ? to_f(xr[c]) : -INFINITY;
    t[j] = ok ? t0[(size_t)c * p.t_sc] : 0.f;
    u[j] = (ok && mix) ? t1[(size_t)c * p.t_sc] : 0.f;
  }
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < CLS_MAXJ; ++j) {
    if (mix) t[j] = lam * t[j] + (1.f - lam) * u[j];
    m = fmaxf(m, x[j]);
  }
  m = warp_max(m);
  if (lane == 0) red[0][warp] = m;
  __syncthreads();
  m = red[0][0];
#pragma unroll
  for (int i = 1; i < NTHREADS / 32; ++i) m = fmaxf(m, red[0][i]);
  __syncthreads();
  float se = 0.f, sum_t = 0.f, sum_tx = 0.f;
#pragma unroll
  for (int j = 0; j < CLS_MAXJ; ++j)
    if (tid + j * NTHREADS < C) {
      se += expf(x[j] - m);
      sum_t += t[j];
      sum_tx = fmaf(t[j], x[j], sum_tx);
    }
  se = warp_sum(se); sum_t = warp_sum(sum_t); sum_tx = warp_sum(sum_tx);
  if (lane == 0) { red[0][warp] = se; red[1][warp] = sum_t; red[2][warp] = sum_tx; }
  __syncthreads();
  se = sum_t = sum_tx = 0.f;
#pragma unroll
  for (int i = 0; i < NTHREADS / 32; ++i) { se += red[0][i]; sum_t += red[1][i]; sum_tx += red[2][i]; }     // fixed order
  const float lse = m + logf(se);
#pragma unroll
  for (int j = 0; j < CLS_MAXJ; ++j) {
    const int c = tid + j * NTHREADS;
    if (c < C) dr[c] = from_f<T>(p.w_cls * (expf(x[j] - lse) * sum_t - t[j]));
  }
  if (tid == 0) p.partial[partial_index] = p.w_cls * (lse * sum_t - sum_tx);
}
template <typename T>
__global__ void __launch_bounds__(NTHREADS) tlce_cls_kernel(TlceParams p, int partial_base) {
  tlce_cls_body<T>(p, blockIdx.x, partial_base + blockIdx.x);
}

// last-CTA reduction of the per-CTA partials: same order as tlce_reduce_kernel (strided per-thread double sums, then a
// fixed tree), so the loss does not depend on which CTA happens to finish last
__device__ __forceinline__ void tlce_ticket_reduce(const TlceParams& p, double* sred) {
  __shared__ int is_last;
  const int tid = threadIdx.x;
  const int total = (int)gridDim.x;
  if (tid == 0) {
    __threadfence();                                    // this CTA's partial is visible before the ticket is taken
    is_last = (atomicAdd(p.ticket, 1) == total - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc = 0.0;
  for (int i = tid; i < total; i += NTHREADS) acc += (double)__ldcg(p.partial + i);
  sred[tid] = acc;
  __syncthreads();
  for (int o = NTHREADS / 2; o > 0; o >>= 1) {
    if (tid < o) sred[tid] += sred[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    *p.loss = (float)sred[0];
    *p.ticket = 0;                                      // re-armed for the next call
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Fast dense path (3-D class-major target, C a multiple of the 16-byte vector width, C <= 1024: ImageNet's C = 1000).
// CTA = 16 tokens of one image, 8 warps, 2 tokens per warp:
//   * each warp first issues the 16-byte loads of ITS two logit rows (4 / 8 vectors per lane, kept in registers), so they
//     are in flight while
//   * all warps transpose the [C, 16] target tile into shared memory as token-major fp32 rows: a warp-load covers 4 class
//     rows x 64 contiguous bytes (8 lanes x float2), and the scalar stores are conflict-free thanks to an XOR swizzle
//     of the 16-byte chunk index with the token-pair index;
//   * then a warp owns a token: max, ONE exp per element (kept in the logit registers), sum_t / sum_tx against 16-byte
//     shared-memory reads of the target row, and the gradient written as 16-byte vectors.
// One HBM read of each logit and target element, one write of each gradient element; 64 KB of shared memory per CTA.
constexpr int FT = 16;          // tokens per CTA
constexpr int FCS = 1024;       // padded row length (floats) of the staged target tile

template <typename T, int MODE> struct FastVec;
template <int MODE> struct FastVec<bf16, MODE> { static constexpr int V = 8, NJ = 4; };
template <int MODE> struct FastVec<float, MODE> { static constexpr int V = 4, NJ = 8; };

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 2) tlce_fast_kernel(TlceParams p) {
  constexpr int V = FastVec<T, 0>::V, NJ = FastVec<T, 0>::NJ;       // elements per 16-byte vector, vectors per lane per token
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* st = reinterpret_cast<float*>(smem_raw);                  // [FT][FCS], 16-byte chunk index XOR ((token >> 1) & 7)
  __shared__ float s_loss[NTHREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, N = p.N;
  // single-launch mode: the B class-token CTAs come FIRST in the grid (their strided label loads are pure latency: started
  // early they overlap the dense tiles instead of forming the tail), the dense tiles follow
  const int cls_ctas = p.ticket != nullptr ? p.B : 0;
  if ((int)blockIdx.x < cls_ctas) {
    tlce_cls_body<T>(p, (int)blockIdx.x, (int)blockIdx.x);
    tlce_ticket_reduce(p, reinterpret_cast<double*>(smem_raw));
    return;
  }
  const int tile = (int)blockIdx.x - cls_ctas;
  const int b = tile / p.tiles_per_img;
  const int n0 = (tile % p.tiles_per_img) * FT;
  const int nt = min(FT, N - n0);
  // ---- (1) this warp's two logit rows -> registers, still PACKED (bf16: 32 registers instead of 64), so that the staging
  //      loop below can keep twice as many target loads in flight under the 128-register budget of two CTAs per SM
  Vec16<T> xraw[2][NJ];
  bool have[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int n = warp * 2 + q;
    have[q] = n < nt;
    const T* xr = reinterpret_cast<const T*>(p.x_aux) + ((size_t)b * N + n0 + n) * C;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = (lane + 32 * j) * V;
      if (have[q] && c < C) xraw[q][j].load(xr + c);
      else xraw[q][j].zero();
    }
  }
  // ---- (2) target tile [C, FT] -> shared memory [FT][FCS]
  {
    const int pq = lane & 7, r = lane >> 3;                          // token pair, class within the warp's group of 4
    const float* tg = p.target + (size_t)b * p.t_sb + (size_t)(p.slot_aux0 + n0 + 2 * pq);
    const bool pair_ok = 2 * pq + 1 < nt, first_ok = 2 * pq < nt;
    const bool vec_ok = (((size_t)b * p.t_sb + (size_t)(p.slot_aux0 + n0)) % 2 == 0) && (p.t_sc % 2 == 0);   // 8-byte aligned float2 loads
    // UNR independent class rows per lane in flight: all loads first, then the swizzled stores (the loop is otherwise a
    // chain of dependent load -> store pairs and runs at DRAM latency, not bandwidth)
    constexpr int UNR = sizeof(T) == 2 ? 16 : 8;
    constexpr int CSTEP = (NTHREADS / 32) * 4;
    for (int cb = warp * 4; cb < C; cb += CSTEP * UNR) {
      float t0[UNR], t1[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int c = cb + u * CSTEP + r;
        t0[u] = 0.f; t1[u] = 0.f;
        if (c < C) {
          const float* src = tg + (size_t)c * p.t_sc;
          if (vec_ok && pair_ok) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(src));
            t0[u] = v.x; t1[u] = v.y;
          } else {
            if (first_ok) t0[u] = __ldg(src);
            if (pair_ok) t1[u] = __ldg(src + 1);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int c = cb + u * CSTEP + r;
        if (c < C) {
          const int pc = c ^ (pq << 2);                               // swizzled column
          st[(2 * pq) * FCS + pc] = t0[u];
          st[(2 * pq + 1) * FCS + pc] = t1[u];
        }
      }
    }
  }
  __syncthreads();
  float x[2][NJ * V];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const bool ok = have[q] && (lane + 32 * j) * V < C;
#pragma unroll
      for (int k = 0; k < V; ++k) x[q][j * V + k] = ok ? xraw[q][j].get(k) : -INFINITY;
    }
  // ---- (3) one token per warp pass
  float my_loss = 0.f;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (!have[q]) continue;
    const int n = warp * 2 + q;
    const float* trow = st + n * FCS;
    const int key = (n >> 1) & 7;
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NJ * V; ++i) m = fmaxf(m, x[q][i]);
    m = warp_max(m);
    float se = 0.f, sum_t = 0.f, sum_tx = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = (lane + 32 * j) * V;
      if (c < C) {
#pragma unroll
        for (int h = 0; h < V / 4; ++h) {
          const float4 t4 = *reinterpret_cast<const float4*>(trow + ((((c >> 2) + h) ^ key) << 2));
          const float tq[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float xv = x[q][j * V + h * 4 + k];
            sum_t += tq[k];
            sum_tx = fmaf(tq[k], xv, sum_tx);
            const float e = __expf(xv - m);
            x[q][j * V + h * 4 + k] = e;
            se += e;
          }
        }
      }
    }
    se = warp_sum(se);
    sum_t = warp_sum(sum_t);
    sum_tx = warp_sum(sum_tx);
    const float lse = m + logf(se);
    if (lane == 0) my_loss += p.w_dense * (lse * sum_t - sum_tx);
    const float a = p.w_dense * sum_t / se;                          // w * sum_t * softmax = a * e
    T* dr = reinterpret_cast<T*>(p.d_aux) + ((size_t)b * N + n0 + n) * C;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = (lane + 32 * j) * V;
      if (c < C) {
        Vec16<T> v;
#pragma unroll
        for (int h = 0; h < V / 4; ++h) {
          const float4 t4 = *reinterpret_cast<const float4*>(trow + ((((c >> 2) + h) ^ key) << 2));     // second shared read, no registers held
          v.set(h * 4 + 0, fmaf(a, x[q][j * V + h * 4 + 0], -p.w_dense * t4.x));
          v.set(h * 4 + 1, fmaf(a, x[q][j * V + h * 4 + 1], -p.w_dense * t4.y));
          v.set(h * 4 + 2, fmaf(a, x[q][j * V + h * 4 + 2], -p.w_dense * t4.z));
          v.set(h * 4 + 3, fmaf(a, x[q][j * V + h * 4 + 3], -p.w_dense * t4.w));
        }
        v.store(dr + c);
      }
    }
  }
  if (lane == 0) s_loss[warp] = my_loss;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < NTHREADS / 32; ++i) s += s_loss[i];
    p.partial[blockIdx.x] = s;
  }
  if (p.ticket != nullptr) {
    __syncthreads();                                               // every warp is done with the staged tile
    tlce_ticket_reduce(p, reinterpret_cast<double*>(smem_raw));
  }
}

// deterministic final reduction of the per-CTA partials (fixed tree order)
__global__ void __launch_bounds__(256) tlce_reduce_kernel(const float* __restrict__ partial, int n, float* loss) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += (double)partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)s[0];
}

template <typename T>
__global__ void scale_by_scalar_kernel(const T* __restrict__ in, T* __restrict__ out, size_t n, const float* scalar) {
  const float s = *scalar;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = from_f<T>(to_f(in[i]) * s);
}

}  // namespace

// workspace: floats, at least apb_tlce_workspace_floats(B, N) entries.
long long apb_tlce_workspace_floats(int B, int N) {
  return (long long)B * ((N + TT - 1) / TT) + B;      // dense tiles + one class-token partial per image (fast path)
}

int apb_tlce_fwd_bwd(const void* x_cls, const void* x_aux, const float* target, int target_is_3d, int B, int N, int C,
                     int box_area, const int* box_dev, float w_cls, float w_dense, float* loss, void* d_cls, void* d_aux,
                     float* workspace, int* ticket, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(B > 0 && N > 0 && C > 0, APB_ERR_SHAPE, "tlce: bad shape B=%d N=%d C=%d", B, N, C);
  APB_CHECK_ARG(dtype == APB_F32 || dtype == APB_BF16, APB_ERR_DTYPE, "tlce: dtype %d", dtype);
  TlceParams p;
  p.x_cls = x_cls; p.x_aux = x_aux; p.target = target; p.d_cls = d_cls; p.d_aux = d_aux; p.partial = workspace;
  p.B = B; p.N = N; p.C = C;
  if (target_is_3d) { p.t_sb = (long long)C * (2 + N); p.t_sc = 2 + N; p.t_ss = 1; p.slot_cls = 1; p.slot_aux0 = 2; }
  else { p.t_sb = C; p.t_sc = 1; p.t_ss = 0; p.slot_cls = 0; p.slot_aux0 = 0; }
  p.lam = 1.f - (float)((double)box_area / (double)N);
  p.box_dev = box_dev;
  p.w_cls = w_cls / (float)B;
  p.w_dense = w_dense / ((float)B * (float)N);
  p.tiles_per_img = (N + TT - 1) / TT;
  const int n_cls_ctas = (B + NTHREADS / 32 - 1) / (NTHREADS / 32);
  const int grid = B * p.tiles_per_img + n_cls_ctas;
  const size_t esz = dtype == APB_F32 ? 4 : 2;
  p.cta_offset = 0;
  p.ticket = nullptr; p.loss = loss; p.n_aux = B * p.tiles_per_img;
  cudaError_t e;
  // fast dense path: class-major 3-D target, 16-byte logit rows, C <= 1024
  const int vecw = dtype == APB_F32 ? 4 : 8;
  const bool fast = target_is_3d && C % vecw == 0 && C <= FCS && (((uintptr_t)x_aux | (uintptr_t)d_aux) & 15) == 0;
  if (fast) {
    const size_t fsmem = (size_t)FT * FCS * sizeof(float);
    const int n_aux = B * p.tiles_per_img;
    p.ticket = ticket;                                   // non-null: ONE launch (dense tiles + class tokens + last-CTA reduce)
    const int fgrid = ticket != nullptr ? n_aux + B : n_aux;
    if (dtype == APB_F32) {
      e = cudaFuncSetAttribute(tlce_fast_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
      if (e != cudaSuccess) { apb_set_error("tlce: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
      tlce_fast_kernel<float><<<fgrid, NTHREADS, fsmem, st>>>(p);
    } else {
      e = cudaFuncSetAttribute(tlce_fast_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
      if (e != cudaSuccess) { apb_set_error("tlce: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
      tlce_fast_kernel<bf16><<<fgrid, NTHREADS, fsmem, st>>>(p);
    }
    APB_LAUNCH_CHECK("tlce_fast_kernel");
    if (ticket != nullptr) return 0;
    if (dtype == APB_F32) tlce_cls_kernel<float><<<B, NTHREADS, 0, st>>>(p, n_aux);
    else tlce_cls_kernel<bf16><<<B, NTHREADS, 0, st>>>(p, n_aux);
    APB_LAUNCH_CHECK("tlce_cls_kernel");
    tlce_reduce_kernel<<<1, 256, 0, st>>>(workspace, n_aux + B, loss);
    APB_LAUNCH_CHECK("tlce_reduce");
    return 0;
  }
  const size_t smem = (size_t)C * TS * sizeof(float) + (size_t)TT * C * esz;
  APB_CHECK_ARG(smem <= 227 * 1024, APB_ERR_UNSUPPORTED, "tlce: C=%d needs %zu B of shared memory (> 227 KB)", C, smem);
  if (dtype == APB_F32) {
    e = cudaFuncSetAttribute(tlce_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { apb_set_error("tlce: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    tlce_kernel<float><<<grid, NTHREADS, smem, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(tlce_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { apb_set_error("tlce: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    tlce_kernel<bf16><<<grid, NTHREADS, smem, st>>>(p);
  }
  APB_LAUNCH_CHECK("tlce_kernel");
  tlce_reduce_kernel<<<1, 256, 0, st>>>(workspace, grid, loss);
  APB_LAUNCH_CHECK("tlce_reduce");
  return 0;
}

// Lazy in-place scaling for the loss backward: buf *= g / applied, skipped entirely (no memory traffic) when the factor is
// exactly 1 -- the case of every training step, where the loss is the root of backward.  `applied` (device float, starts
// at 1) remembers the factor already folded into buf, so repeated backward passes with other upstream gradients stay
// correct.  Two launches: the scaling kernel only READS applied; a one-thread kernel then records the new value.
template <typename T>
__global__ void scale_lazy_kernel(T* __restrict__ buf, size_t n, const float* __restrict__ g, const float* __restrict__ applied) {
  const float f = *g / *applied;
  if (f == 1.f) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    buf[i] = from_f<T>(to_f(buf[i]) * f);
}
__global__ void scale_lazy_commit_kernel(const float* __restrict__ g, float* __restrict__ applied) { *applied = *g; }

int apb_scale_lazy(void* buf_a, long long n_a, void* buf_b, long long n_b, const float* g, float* applied, int dtype,
                   apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  for (int i = 0; i < 2; ++i) {
    void* buf = i ? buf_b : buf_a;
    const long long n = i ? n_b : n_a;
    if (buf == nullptr || n <= 0) continue;
    const int grid = (int)((n + 1023) / 1024 > 148 * 8 ? 148 * 8 : (n + 1023) / 1024);
    if (dtype == APB_F32) scale_lazy_kernel<float><<<grid, 256, 0, st>>>((float*)buf, (size_t)n, g, applied);
    else scale_lazy_kernel<bf16><<<grid, 256, 0, st>>>((bf16*)buf, (size_t)n, g, applied);
    APB_LAUNCH_CHECK("scale_lazy");
  }
  scale_lazy_commit_kernel<<<1, 1, 0, st>>>(g, applied);
  APB_LAUNCH_CHECK("scale_lazy_commit");
  return 0;
}

// out = in * (*scalar)  (backward of the fused loss: grads were produced for upstream gradient 1)
int apb_scale_by_scalar(const void* in, void* out, long long n, const float* scalar, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (n <= 0) return 0;
  const int grid = (int)((n + 1023) / 1024 > 148 * 8 ? 148 * 8 : (n + 1023) / 1024);
  if (dtype == APB_F32) scale_by_scalar_kernel<float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, (size_t)n, scalar);
  else scale_by_scalar_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)in, (bf16*)out, (size_t)n, scalar);
  APB_LAUNCH_CHECK("scale_by_scalar");
  return 0;
}
