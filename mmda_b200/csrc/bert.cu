// BERT-base text encoder pieces: the kernels that, together with the
// tcgen05 GEMMs (gemm_tc.cu), LayerNorm and dropout (elementwise.cu), replace
//   BertModel(input_ids, attention_mask, token_type_ids)[0]  ->  masked mean
// at reference src/models.py:41-45,186-198 (HF `BertModel`, bert-base geometry: 12 layers, hidden
// 768, 12 heads of 64, intermediate 3072, erf-GELU, LayerNorm eps 1e-12, dropout 0.1 on the
// embeddings, the attention probabilities and the two dense outputs of every layer).
//
// Layout: tokens are batch-first, row m = b*S + s of [B*S][hidden]; QKV rows are
// [q(hidden) | k(hidden) | v(hidden)]; probabilities [B][heads][S][S] (post-softmax, pre-dropout).
#include "common.cuh"
#include <cuda_bf16.h>

// ---------------------------------------------------------------- embeddings ---------------
// BertEmbeddings: word[ids] + position[s] + token_type[types]   (then LayerNorm + dropout)
__global__ void bert_embed_fwd_kernel(const float* __restrict__ word, const float* __restrict__ pos,
                                      const float* __restrict__ typ,
                                      const long long* __restrict__ ids,
                                      const long long* __restrict__ types, int M, int S, int H,
                                      int V, float* __restrict__ out) {
  const int m = blockIdx.x;
  if (m >= M) return;
  long long id = ids[m];
  if (id < 0 || id >= V) id = 0;   // torch would raise; ids are validated on the host
  const long long ty = types[m] != 0 ? 1 : 0;
  const float* w = word + (size_t)id * H;
  const float* p = pos + (size_t)(m % S) * H;
  const float* t = typ + (size_t)ty * H;
  float* o = out + (size_t)m * H;
  for (int c = threadIdx.x; c < H; c += blockDim.x) o[c] = w[c] + p[c] + t[c];
}

// word_embeddings has padding_idx = 0: its row never receives gradient
__global__ void bert_embed_bwd_kernel(const float* __restrict__ d, const long long* __restrict__ ids,
                                      const long long* __restrict__ types, int M, int S, int H,
                                      int V, float* __restrict__ dword, float* __restrict__ dpos,
                                      float* __restrict__ dtyp) {
  const int m = blockIdx.x;
  if (m >= M) return;
  const long long id = ids[m];
  const long long ty = types[m] != 0 ? 1 : 0;
  const float* g = d + (size_t)m * H;
  const bool w_ok = dword != nullptr && id > 0 && id < V;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float v = g[c];
    if (v == 0.f) continue;        // padded positions carry exact zeros
    if (w_ok) atomicAdd(dword + (size_t)id * H + c, v);
    if (dpos) atomicAdd(dpos + (size_t)(m % S) * H + c, v);
    if (dtyp) atomicAdd(dtyp + (size_t)ty * H + c, v);
  }
}

// ---------------------------------------------------------------- GELU (erf form) ----------
__device__ __forceinline__ float gelu_f(float v) {
  return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad_f(float v) {
  const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * expf(-0.5f * v * v);
  return cdf + v * pdf;
}
// n4 = n / 4 float4 groups.  y / dx may be NULL when only the bf16 operand copy (ybf / dxbf) is
// wanted: in bf16 mode the GELU output and its gradient are consumed by tensor-core GEMMs only.
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&lo);
  pk.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                __nv_bfloat16* __restrict__ ybf, size_t n4, size_t n) {
  const size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = t0; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 r = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
    if (y) reinterpret_cast<float4*>(y)[i] = r;
    if (ybf) st_bf16x4(ybf + i * 4, r.x, r.y, r.z, r.w);
  }
  for (size_t i = n4 * 4 + t0; i < n; i += stride) {
    const float r = gelu_f(x[i]);
    if (y) y[i] = r;
    if (ybf) ybf[i] = __float2bfloat16(r);
  }
}
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                float* __restrict__ dx, __nv_bfloat16* __restrict__ dxbf, size_t n4,
                                size_t n) {
  const size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = t0; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 d = reinterpret_cast<const float4*>(dy)[i];
    const float4 r = make_float4(d.x * gelu_grad_f(v.x), d.y * gelu_grad_f(v.y),
                                 d.z * gelu_grad_f(v.z), d.w * gelu_grad_f(v.w));
    if (dx) reinterpret_cast<float4*>(dx)[i] = r;
    if (dxbf) st_bf16x4(dxbf + i * 4, r.x, r.y, r.z, r.w);
  }
  for (size_t i = n4 * 4 + t0; i < n; i += stride) {
    const float r = dy[i] * gelu_grad_f(x[i]);
    if (dx) dx[i] = r;
    if (dxbf) dxbf[i] = __float2bfloat16(r);
  }
}

// ---------------------------------------------------------------- masked mean --------------
// utt[b] = sum_s mask[b,s] * hid[b,s] / sum_s mask[b,s]          (src/models.py:192-196)
__global__ void masked_mean_fwd_kernel(const float* __restrict__ hid,
                                       const long long* __restrict__ mask, int S, int H,
                                       float* __restrict__ utt) {
  const int b = blockIdx.x;
  float cnt = 0.f;
  for (int s = 0; s < S; ++s) cnt += (float)mask[(size_t)b * S + s];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s)
      acc += (float)mask[(size_t)b * S + s] * hid[((size_t)b * S + s) * H + c];
    utt[(size_t)b * H + c] = acc / cnt;
  }
}
__global__ void masked_mean_bwd_kernel(const float* __restrict__ dutt,
                                       const long long* __restrict__ mask, int S, int H,
                                       float* __restrict__ dhid) {
  const int m = blockIdx.x, b = m / S;
  float cnt = 0.f;
  for (int s = 0; s < S; ++s) cnt += (float)mask[(size_t)b * S + s];
  const float w = (float)mask[m] / cnt;
  for (int c = threadIdx.x; c < H; c += blockDim.x)
    dhid[(size_t)m * H + c] = w * dutt[(size_t)b * H + c];
}

// ---------------------------------------------------------------- self-attention -----------
// One CTA per (sample, head); the whole S x S problem lives in shared memory (S <= ~100).
// Every contraction is a register-tiled 4x4 outer product over "k-major" operands
// (At[k][I], Bt[k][J], leading dimensions multiples of 4, zero padded): two LDS.128 feed 16 FMAs,
// so the kernels are FMA-issue bound rather than shared-memory bound.
constexpr int BHD = 64;          // head dim of bert-base
constexpr int ATT_THREADS = 256;

__device__ __forceinline__ float drop_scale(unsigned long long seed, unsigned stream, unsigned idx,
                                            float p, float inv_keep) {
  return (p > 0.f && rng_uniform(seed, stream, idx) < p) ? 0.f : inv_keep;
}

// acc[a][b] = sum_k At[k*lda + i0 + a] * Bt[k*ldb + j0 + b]
__device__ __forceinline__ void mm_tile(const float* __restrict__ At, int lda,
                                        const float* __restrict__ Bt, int ldb, int K, int i0, int j0,
                                        float (&acc)[4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const float* ap = At + i0;
  const float* bp = Bt + j0;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 av = *reinterpret_cast<const float4*>(ap + (size_t)k * lda);
    const float4 bv = *reinterpret_cast<const float4*>(bp + (size_t)k * ldb);
    const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
bert_attn_fwd_kernel(const float* __restrict__ qkv, const long long* __restrict__ mask,
                     float* __restrict__ ctx, float* __restrict__ probs, int S, int nhead,
                     float scale, float p_drop, unsigned long long seed,
                     const unsigned long long* __restrict__ seed_dev, unsigned stream) {
  extern __shared__ __align__(16) float sm[];
  if (seed_dev) seed += seed_dev[0] * 0x9E3779B97F4A7C15ULL;
  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int Hd = nhead * BHD, ld = 3 * Hd;
  const int S4 = (S + 3) & ~3;
  float* Qt = sm;                  // [64][S4]
  float* Kt = Qt + BHD * S4;       // [64][S4]
  float* Vs = Kt + BHD * S4;       // [S4][64]
  float* Sc = Vs + S4 * BHD;       // [S4][S4] scores -> probabilities
  float* Pt = Sc + S4 * S4;        // [S4][S4] dropped probabilities, transposed: Pt[j][i]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  // 16-byte global loads, several in flight per thread.  Lanes run over the token index so the
  // transposed (k-major) stores are bank-conflict free; the 16-byte row segments each lane reads
  // are served from L2.
#pragma unroll 4
  for (int idx = tid; idx < S4 * (BHD / 4); idx += blockDim.x) {
    const int s = idx % S4, c = (idx / S4) * 4;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q, v = q;
    if (s < S) {
      const float* r = qkv + (size_t)(b * S + s) * ld + h * BHD + c;
      q = *reinterpret_cast<const float4*>(r);
      k = *reinterpret_cast<const float4*>(r + Hd);
      v = *reinterpret_cast<const float4*>(r + 2 * Hd);
    }
    Qt[(c + 0) * S4 + s] = q.x; Qt[(c + 1) * S4 + s] = q.y; Qt[(c + 2) * S4 + s] = q.z; Qt[(c + 3) * S4 + s] = q.w;
    Kt[(c + 0) * S4 + s] = k.x; Kt[(c + 1) * S4 + s] = k.y; Kt[(c + 2) * S4 + s] = k.z; Kt[(c + 3) * S4 + s] = k.w;
    *reinterpret_cast<float4*>(Vs + s * BHD + c) = v;
  }
  for (int idx = tid; idx < S4 * S4; idx += blockDim.x) Pt[idx] = 0.f;
  __syncthreads();
  const int nb = S4 >> 2;
  for (int blk = tid; blk < nb * nb; blk += blockDim.x) {
    const int i0 = (blk / nb) * 4, j0 = (blk % nb) * 4;
    float acc[4][4];
    mm_tile(Qt, S4, Kt, S4, BHD, i0, j0, acc);
    bool keep[4];      // HF adds finfo.min to masked keys: the softmax weight is exactly 0
#pragma unroll
    for (int c = 0; c < 4; ++c) keep[c] = (j0 + c < S) && mask[(size_t)b * S + j0 + c] != 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
      *reinterpret_cast<float4*>(Sc + (i0 + a) * S4 + j0) =
          make_float4(keep[0] ? acc[a][0] * scale : -INFINITY, keep[1] ? acc[a][1] * scale : -INFINITY,
                      keep[2] ? acc[a][2] * scale : -INFINITY, keep[3] ? acc[a][3] * scale : -INFINITY);
  }
  __syncthreads();
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int i = warp; i < S; i += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, Sc[i * S4 + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float e = expf(Sc[i * S4 + j] - mx);
      Sc[i * S4 + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    const unsigned base = (unsigned)(((b * nhead + h) * S + i) * S);
    for (int j = lane; j < S; j += 32) {
      const float pr = Sc[i * S4 + j] * inv;
      if (probs) probs[(size_t)base + j] = pr;
      Pt[j * S4 + i] = pr * drop_scale(seed, stream, base + j, p_drop, inv_keep);
    }
  }
  __syncthreads();
  for (int blk = tid; blk < nb * (BHD / 4); blk += blockDim.x) {
    const int i0 = (blk / (BHD / 4)) * 4, c0 = (blk % (BHD / 4)) * 4;
    float acc[4][4];
    mm_tile(Pt, S4, Vs, BHD, S, i0, c0, acc);
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (i0 + a < S)
        *reinterpret_cast<float4*>(ctx + (size_t)(b * S + i0 + a) * Hd + h * BHD + c0) =
            make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
bert_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                     const float* __restrict__ dctx, float* __restrict__ dqkv, int S, int nhead,
                     float scale, float p_drop, unsigned long long seed,
                     const unsigned long long* __restrict__ seed_dev, unsigned stream) {
  extern __shared__ __align__(16) float sm[];
  if (seed_dev) seed += seed_dev[0] * 0x9E3779B97F4A7C15ULL;
  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int Hd = nhead * BHD, ld = 3 * Hd;
  const int S4 = (S + 3) & ~3;
  float* Qs = sm;                  // [S4][64]
  float* Ks = Qs + S4 * BHD;       // [S4][64]
  float* Vt = Ks + S4 * BHD;       // [64][S4]
  float* Cs = Vt + BHD * S4;       // d(ctx) [S4][64]
  float* Ct = Cs + S4 * BHD;       // d(ctx) transposed [64][S4]
  float* Ps = Ct + BHD * S4;       // [S4][S4] probabilities -> dropped probabilities (row major)
  float* Ds = Ps + S4 * S4;        // [S4][S4] d(probs) -> d(scores) (row major)
  float* Dt = Ds + S4 * S4;        // d(scores) transposed
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const unsigned base0 = (unsigned)((b * nhead + h) * S * S);
#pragma unroll 2
  for (int idx = tid; idx < S4 * (BHD / 4); idx += blockDim.x) {
    const int s = idx % S4, c = (idx / S4) * 4;      // lanes over tokens: conflict-free transposes
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q, v = q, dc = q;
    if (s < S) {
      const float* r = qkv + (size_t)(b * S + s) * ld + h * BHD + c;
      q = *reinterpret_cast<const float4*>(r);
      k = *reinterpret_cast<const float4*>(r + Hd);
      v = *reinterpret_cast<const float4*>(r + 2 * Hd);
      dc = *reinterpret_cast<const float4*>(dctx + (size_t)(b * S + s) * Hd + h * BHD + c);
    }
    *reinterpret_cast<float4*>(Qs + s * BHD + c) = q;
    *reinterpret_cast<float4*>(Ks + s * BHD + c) = k;
    *reinterpret_cast<float4*>(Cs + s * BHD + c) = dc;
    Vt[(c + 0) * S4 + s] = v.x; Vt[(c + 1) * S4 + s] = v.y; Vt[(c + 2) * S4 + s] = v.z; Vt[(c + 3) * S4 + s] = v.w;
    Ct[(c + 0) * S4 + s] = dc.x; Ct[(c + 1) * S4 + s] = dc.y; Ct[(c + 2) * S4 + s] = dc.z; Ct[(c + 3) * S4 + s] = dc.w;
  }
#pragma unroll 4
  for (int idx = tid; idx < S4 * S4; idx += blockDim.x) {
    const int i = idx / S4, j = idx % S4;
    Ps[idx] = (i < S && j < S) ? probs[(size_t)base0 + i * S + j] : 0.f;
    Dt[idx] = 0.f;
  }
  __syncthreads();
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const int nb = S4 >> 2;
  for (int blk = tid; blk < nb * nb; blk += blockDim.x) {      // d(dropped probs) -> d(probs)
    const int i0 = (blk / nb) * 4, j0 = (blk % nb) * 4;
    float acc[4][4];
    mm_tile(Ct, S4, Vt, S4, BHD, i0, j0, acc);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + a;
      float o[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + c;
        o[c] = (i < S && j < S)
            ? acc[a][c] * drop_scale(seed, stream, base0 + i * S + j, p_drop, inv_keep) : 0.f;
      }
      *reinterpret_cast<float4*>(Ds + i * S4 + j0) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();
  for (int i = warp; i < S; i += nw) {      // softmax backward; probabilities -> dropped ones
    float r = 0.f;
    for (int j = lane; j < S; j += 32) r += Ds[i * S4 + j] * Ps[i * S4 + j];
    r = warp_sum(r);
    for (int j = lane; j < S; j += 32) {
      const float pr = Ps[i * S4 + j];
      const float ds = pr * (Ds[i * S4 + j] - r) * scale;
      Ds[i * S4 + j] = ds;
      Dt[j * S4 + i] = ds;
      Ps[i * S4 + j] = pr * drop_scale(seed, stream, base0 + i * S + j, p_drop, inv_keep);
    }
  }
  __syncthreads();
  const int cb = BHD / 4;
  for (int blk = tid; blk < 3 * nb * cb; blk += blockDim.x) {
    const int which = blk / (nb * cb), rem = blk % (nb * cb);
    const int s0 = (rem / cb) * 4, c0 = (rem % cb) * 4;
    float acc[4][4];
    if (which == 0) mm_tile(Dt, S4, Ks, BHD, S, s0, c0, acc);        // dQ[i] = sum_j dS[i][j] K[j]
    else if (which == 1) mm_tile(Ds, S4, Qs, BHD, S, s0, c0, acc);   // dK[j] = sum_i dS[i][j] Q[i]
    else mm_tile(Ps, S4, Cs, BHD, S, s0, c0, acc);                   // dV[j] = sum_i Pd[i][j] dC[i]
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (s0 + a < S)
        *reinterpret_cast<float4*>(dqkv + (size_t)(b * S + s0 + a) * ld + which * Hd + h * BHD + c0) =
            make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
  }
}


// ------------------------------------------------------------------------------------------
// bf16-mode attention core on the tensor pipe (BASELINE configs[3] runs the encoder in bf16):
// mma.sync m16n8k16 bf16 with fp32 accumulation, one (sample, head) per 4-warp CTA, the whole
// sequence (S <= 64) resident.  Forward is flash-style: the 16 x 64 score fragment of each warp
// stays in registers through mask -> softmax -> dropout and is re-used as the A operand of P*V.
// The backward needs dS and the dropped probabilities transposed, so those two go through shared
// memory once (bf16); everything else is fragments.  Same dropout stream / element indexing as
// the fp32 kernels above, so either backward matches either forward.  tcgen05 is the wrong tool
// here: a 52 x 52 x 64 problem per head is a quarter of one 128-row UMMA tile.
// ------------------------------------------------------------------------------------------
constexpr int AP = 72;            // smem row pitch, bf16 elements: 144 B keeps ldmatrix conflict-free
constexpr int ATT_MMA_S = 64;     // resident sequence length

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __nv_bfloat16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// fragment addresses (lane -> the 8x8 matrix row it points ldmatrix at)
//   A, smem X[m][k]                 : &X[(m0 + (lane & 15)) * AP + k0 + (lane >> 4) * 8]
//   A, smem W[k][m]   (.trans)      : &W[(k0 + (q >> 1) * 8 + r) * AP + m0 + (q & 1) * 8]
//   B, smem Y[n][k]   (two n-tiles) : &Y[(n0 + (q >> 1) * 8 + r) * AP + k0 + (q & 1) * 8]
//   B, smem Z[k][n]   (.trans, two) : &Z[(k0 + (q & 1) * 8 + r) * AP + n0 + (q >> 1) * 8]
// with q = lane >> 3, r = lane & 7; for B the registers come back as {b0, b1} of n-tile n0 and
// {b0, b1} of n-tile n0 + 8.

// stage rows [0, 64) x 64 features of a [token][ld] fp32 matrix as bf16 (rows >= S zero)
__device__ __forceinline__ void stage_bf16(__nv_bfloat16* dst, const float* src, int ld, int S, int tid) {
#pragma unroll 4
  for (int idx = tid; idx < ATT_MMA_S * 16; idx += 128) {
    const int s = idx >> 4, c = (idx & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < S) v = *reinterpret_cast<const float4*>(src + (size_t)s * ld + c);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + s * AP + c) = o;
  }
}

__global__ void __launch_bounds__(128)
bert_attn_fwd_mma_kernel(const float* __restrict__ qkv, const long long* __restrict__ mask,
                         float* __restrict__ ctx, __nv_bfloat16* __restrict__ ctx_bf16,
                         float* __restrict__ probs, int S, int nhead,
                         float scale, float p_drop, unsigned long long seed,
                         const unsigned long long* __restrict__ seed_dev, unsigned stream) {
  __shared__ __align__(16) __nv_bfloat16 Qs[ATT_MMA_S * AP], Ks[ATT_MMA_S * AP], Vs[ATT_MMA_S * AP];
  __shared__ float keep_s[ATT_MMA_S];
  if (seed_dev) seed += seed_dev[0] * 0x9E3779B97F4A7C15ULL;
  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int Hd = nhead * BHD, ld = 3 * Hd;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* base = qkv + (size_t)b * S * ld + h * BHD;
  stage_bf16(Qs, base, ld, S, tid);
  stage_bf16(Ks, base + Hd, ld, S, tid);
  stage_bf16(Vs, base + 2 * Hd, ld, S, tid);
  if (tid < ATT_MMA_S) keep_s[tid] = (tid < S && mask[(size_t)b * S + tid] != 0) ? 1.f : 0.f;
  __syncthreads();
  const int m0 = warp * 16;
  if (m0 >= S) return;                       // no block-wide barrier below
  const int q = lane >> 3, r = lane & 7, g = lane >> 2, t4 = lane & 3;
  float sc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {           // scores = Q K^T
    uint32_t a[4];
    ldsm_x4(a, &Qs[(m0 + (lane & 15)) * AP + kt * 16 + (lane >> 4) * 8]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bk[4];
      ldsm_x4(bk, &Ks[(np * 16 + (q >> 1) * 8 + r) * AP + kt * 16 + (q & 1) * 8]);
      mma_bf16(sc[2 * np], a, bk[0], bk[1]);
      mma_bf16(sc[2 * np + 1], a, bk[2], bk[3]);
    }
  }
  // this thread holds rows i0 = m0 + g (elements 0,1) and i1 = i0 + 8 (elements 2,3), columns
  // j = 8 n + 2 t4 + {0, 1}
  const int i0 = m0 + g, i1 = i0 + 8;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = n * 8 + t4 * 2 + (e & 1);
      // HF adds finfo.min to masked keys: the softmax weight is exactly 0
      const float v = keep_s[j] != 0.f ? sc[n][e] * scale : -INFINITY;
      sc[n][e] = v;
      if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
    }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ex = expf(sc[n][e] - (e < 2 ? mx0 : mx1));
      sc[n][e] = ex;
      if (e < 2) s0 += ex; else s1 += ex;
    }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float inv0 = 1.f / s0, inv1 = 1.f / s1;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const unsigned base0 = (unsigned)((b * nhead + h) * S * S);
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = e < 2 ? i0 : i1, j = n * 8 + t4 * 2 + (e & 1);
      const float pr = sc[n][e] * (e < 2 ? inv0 : inv1);
      float pd = 0.f;
      if (i < S && j < S) {
        if (probs) probs[(size_t)base0 + i * S + j] = pr;
        pd = pr * drop_scale(seed, stream, base0 + i * S + j, p_drop, inv_keep);
      }
      sc[n][e] = pd;
    }
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {           // context = dropped probabilities * V
    uint32_t a[4];
    a[0] = pack_bf16(sc[2 * kt][0], sc[2 * kt][1]);
    a[1] = pack_bf16(sc[2 * kt][2], sc[2 * kt][3]);
    a[2] = pack_bf16(sc[2 * kt + 1][0], sc[2 * kt + 1][1]);
    a[3] = pack_bf16(sc[2 * kt + 1][2], sc[2 * kt + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bv[4];
      ldsm_x4_t(bv, &Vs[(kt * 16 + (q & 1) * 8 + r) * AP + np * 16 + (q >> 1) * 8]);
      mma_bf16(o[2 * np], a, bv[0], bv[1]);
      mma_bf16(o[2 * np + 1], a, bv[2], bv[3]);
    }
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = n * 8 + t4 * 2;
    if (i0 < S) {
      if (ctx) *reinterpret_cast<float2*>(ctx + (size_t)(b * S + i0) * Hd + h * BHD + c) = make_float2(o[n][0], o[n][1]);
      if (ctx_bf16) *reinterpret_cast<uint32_t*>(ctx_bf16 + (size_t)(b * S + i0) * Hd + h * BHD + c) = pack_bf16(o[n][0], o[n][1]);
    }
    if (i1 < S) {
      if (ctx) *reinterpret_cast<float2*>(ctx + (size_t)(b * S + i1) * Hd + h * BHD + c) = make_float2(o[n][2], o[n][3]);
      if (ctx_bf16) *reinterpret_cast<uint32_t*>(ctx_bf16 + (size_t)(b * S + i1) * Hd + h * BHD + c) = pack_bf16(o[n][2], o[n][3]);
    }
  }
}

constexpr int ATT_MMA_BWD_SMEM = 6 * ATT_MMA_S * AP * 2;      // Q K V dO Pd dS, bf16

__global__ void __launch_bounds__(128)
bert_attn_bwd_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                         const float* __restrict__ dctx, float* __restrict__ dqkv,
                         __nv_bfloat16* __restrict__ dqkv_bf16, int S, int nhead,
                         float scale, float p_drop, unsigned long long seed,
                         const unsigned long long* __restrict__ seed_dev, unsigned stream) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* Ks = Qs + ATT_MMA_S * AP;
  __nv_bfloat16* Vs = Ks + ATT_MMA_S * AP;
  __nv_bfloat16* Cs = Vs + ATT_MMA_S * AP;     // d(context)
  __nv_bfloat16* Ps = Cs + ATT_MMA_S * AP;     // dropped probabilities [i][j]
  __nv_bfloat16* Ds = Ps + ATT_MMA_S * AP;     // d(scores) [i][j]
  if (seed_dev) seed += seed_dev[0] * 0x9E3779B97F4A7C15ULL;
  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int Hd = nhead * BHD, ld = 3 * Hd;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* base = qkv + (size_t)b * S * ld + h * BHD;
  stage_bf16(Qs, base, ld, S, tid);
  stage_bf16(Ks, base + Hd, ld, S, tid);
  stage_bf16(Vs, base + 2 * Hd, ld, S, tid);
  stage_bf16(Cs, dctx + (size_t)b * S * Hd + h * BHD, Hd, S, tid);
  __syncthreads();
  const int m0 = warp * 16;
  const int q = lane >> 3, r = lane & 7, g = lane >> 2, t4 = lane & 3;
  const int i0 = m0 + g, i1 = i0 + 8;
  const unsigned base0 = (unsigned)((b * nhead + h) * S * S);
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float dp[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {           // d(dropped probs) = dO V^T
    uint32_t a[4];
    ldsm_x4(a, &Cs[(m0 + (lane & 15)) * AP + kt * 16 + (lane >> 4) * 8]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bv[4];
      ldsm_x4(bv, &Vs[(np * 16 + (q >> 1) * 8 + r) * AP + kt * 16 + (q & 1) * 8]);
      mma_bf16(dp[2 * np], a, bv[0], bv[1]);
      mma_bf16(dp[2 * np + 1], a, bv[2], bv[3]);
    }
  }
  float pr[8][4];
  float r0 = 0.f, r1 = 0.f;
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = e < 2 ? i0 : i1, j = n * 8 + t4 * 2 + (e & 1);
      float p = 0.f, f = 0.f;
      if (i < S && j < S) {
        p = probs[(size_t)base0 + i * S + j];
        f = drop_scale(seed, stream, base0 + i * S + j, p_drop, inv_keep);
      }
      const float d = dp[n][e] * f;          // d(probs)
      dp[n][e] = d;
      pr[n][e] = p;
      if (e < 2) r0 = fmaf(d, p, r0); else r1 = fmaf(d, p, r1);
      Ps[i * AP + j] = __float2bfloat16_rn(p * f);     // dropped probability, read transposed for dV
    }
  r0 += __shfl_xor_sync(0xffffffffu, r0, 1);
  r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
  r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
  r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
#pragma unroll
  for (int n = 0; n < 8; ++n) {              // softmax backward -> d(scores), kept as A fragments
    const int j = n * 8 + t4 * 2;
    dp[n][0] = pr[n][0] * (dp[n][0] - r0) * scale;
    dp[n][1] = pr[n][1] * (dp[n][1] - r0) * scale;
    dp[n][2] = pr[n][2] * (dp[n][2] - r1) * scale;
    dp[n][3] = pr[n][3] * (dp[n][3] - r1) * scale;
    *reinterpret_cast<uint32_t*>(&Ds[i0 * AP + j]) = pack_bf16(dp[n][0], dp[n][1]);
    *reinterpret_cast<uint32_t*>(&Ds[i1 * AP + j]) = pack_bf16(dp[n][2], dp[n][3]);
  }
  float acc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {           // dQ[i] = sum_j dS[i][j] K[j]
    uint32_t a[4];
    a[0] = pack_bf16(dp[2 * kt][0], dp[2 * kt][1]);
    a[1] = pack_bf16(dp[2 * kt][2], dp[2 * kt][3]);
    a[2] = pack_bf16(dp[2 * kt + 1][0], dp[2 * kt + 1][1]);
    a[3] = pack_bf16(dp[2 * kt + 1][2], dp[2 * kt + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bk[4];
      ldsm_x4_t(bk, &Ks[(kt * 16 + (q & 1) * 8 + r) * AP + np * 16 + (q >> 1) * 8]);
      mma_bf16(acc[2 * np], a, bk[0], bk[1]);
      mma_bf16(acc[2 * np + 1], a, bk[2], bk[3]);
    }
  }
  auto store_rows = [&](float (&v)[8][4], int which) {
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int c = n * 8 + t4 * 2;
      const size_t o0 = (size_t)(b * S + i0) * ld + which * Hd + h * BHD + c;
      const size_t o1 = (size_t)(b * S + i1) * ld + which * Hd + h * BHD + c;
      if (i0 < S) {
        if (dqkv) *reinterpret_cast<float2*>(dqkv + o0) = make_float2(v[n][0], v[n][1]);
        if (dqkv_bf16) *reinterpret_cast<uint32_t*>(dqkv_bf16 + o0) = pack_bf16(v[n][0], v[n][1]);
      }
      if (i1 < S) {
        if (dqkv) *reinterpret_cast<float2*>(dqkv + o1) = make_float2(v[n][2], v[n][3]);
        if (dqkv_bf16) *reinterpret_cast<uint32_t*>(dqkv_bf16 + o1) = pack_bf16(v[n][2], v[n][3]);
      }
    }
  };
  store_rows(acc, 0);
  __syncthreads();                           // every warp's rows of Pd / dS are in shared memory
  float dv[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {           // key rows m0..: dK[j] = sum_i dS[i][j] Q[i]; dV[j] = sum_i Pd[i][j] dO[i]
    uint32_t ad[4], ap[4];
    ldsm_x4_t(ad, &Ds[(kt * 16 + (q >> 1) * 8 + r) * AP + m0 + (q & 1) * 8]);
    ldsm_x4_t(ap, &Ps[(kt * 16 + (q >> 1) * 8 + r) * AP + m0 + (q & 1) * 8]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bq[4], bc[4];
      ldsm_x4_t(bq, &Qs[(kt * 16 + (q & 1) * 8 + r) * AP + np * 16 + (q >> 1) * 8]);
      mma_bf16(acc[2 * np], ad, bq[0], bq[1]);
      mma_bf16(acc[2 * np + 1], ad, bq[2], bq[3]);
      ldsm_x4_t(bc, &Cs[(kt * 16 + (q & 1) * 8 + r) * AP + np * 16 + (q >> 1) * 8]);
      mma_bf16(dv[2 * np], ap, bc[0], bc[1]);
      mma_bf16(dv[2 * np + 1], ap, bc[2], bc[3]);
    }
  }
  store_rows(acc, 1);
  store_rows(dv, 2);
}

static inline int ew_grid_b(size_t n) {
  size_t g = (n + 255) / 256;
  return (int)(g > 148 * 16 ? 148 * 16 : (g == 0 ? 1 : g));
}

extern "C" {

int mmda_bert_embed_forward(const float* word, const float* pos, const float* typ,
                            const long long* ids, const long long* types, int B, int S, int H,
                            int V, int max_pos, float* out, cudaStream_t stream) {
  MMDA_REQUIRE(B > 0 && S > 0 && H > 0, "bert_embed: bad sizes B=%d S=%d H=%d", B, S, H);
  MMDA_REQUIRE(S <= max_pos, "bert_embed: sequence %d exceeds the %d position embeddings", S, max_pos);
  bert_embed_fwd_kernel<<<B * S, 256, 0, stream>>>(word, pos, typ, ids, types, B * S, S, H, V, out);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

int mmda_bert_embed_backward(const float* d, const long long* ids, const long long* types, int B,
                             int S, int H, int V, float* dword, float* dpos, float* dtyp,
                             cudaStream_t stream) {
  MMDA_REQUIRE(B > 0 && S > 0 && H > 0, "bert_embed_backward: bad sizes B=%d S=%d H=%d", B, S, H);
  bert_embed_bwd_kernel<<<B * S, 256, 0, stream>>>(d, ids, types, B * S, S, H, V, dword, dpos, dtyp);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

int mmda_gelu_forward(const float* x, float* y, void* y_bf16, long long n, cudaStream_t stream) {
  if (n <= 0) return MMDA_OK;
  MMDA_REQUIRE(y != nullptr || y_bf16 != nullptr, "gelu_forward: no output");
  const bool v4 = (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && ((uintptr_t)y_bf16 & 7) == 0;
  const size_t n4 = v4 ? (size_t)n / 4 : 0;
  gelu_fwd_kernel<<<ew_grid_b(v4 ? n4 + 1 : (size_t)n), 256, 0, stream>>>(
      x, y, reinterpret_cast<__nv_bfloat16*>(y_bf16), n4, (size_t)n);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

int mmda_gelu_backward(const float* dy, const float* x, float* dx, void* dx_bf16, long long n,
                       cudaStream_t stream) {
  if (n <= 0) return MMDA_OK;
  MMDA_REQUIRE(dx != nullptr || dx_bf16 != nullptr, "gelu_backward: no output");
  const bool v4 = (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0 && ((uintptr_t)dx_bf16 & 7) == 0;
  const size_t n4 = v4 ? (size_t)n / 4 : 0;
  gelu_bwd_kernel<<<ew_grid_b(v4 ? n4 + 1 : (size_t)n), 256, 0, stream>>>(
      dy, x, dx, reinterpret_cast<__nv_bfloat16*>(dx_bf16), n4, (size_t)n);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

int mmda_masked_mean_forward(const float* hid, const long long* mask, int B, int S, int H,
                             float* utt, cudaStream_t stream) {
  MMDA_REQUIRE(B > 0 && S > 0 && H > 0, "masked_mean: bad sizes B=%d S=%d H=%d", B, S, H);
  masked_mean_fwd_kernel<<<B, 256, 0, stream>>>(hid, mask, S, H, utt);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

int mmda_masked_mean_backward(const float* dutt, const long long* mask, int B, int S, int H,
                              float* dhid, cudaStream_t stream) {
  MMDA_REQUIRE(B > 0 && S > 0 && H > 0, "masked_mean_backward: bad sizes B=%d S=%d H=%d", B, S, H);
  masked_mean_bwd_kernel<<<B * S, 256, 0, stream>>>(dutt, mask, S, H, dhid);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

static int bert_attn_smem(int S, bool bwd) {
  const int S4 = (S + 3) & ~3;
  return (int)(((bwd ? 5 : 3) * S4 * BHD + (bwd ? 3 : 2) * S4 * S4) * sizeof(float));
}

int mmda_bert_attention_forward(const float* qkv, const long long* mask, float* ctx, float* probs,
                                int B, int S, int nhead, int head_dim, float p_drop,
                                unsigned long long seed, const unsigned long long* seed_dev,
                                unsigned stream_id, cudaStream_t stream) {
  MMDA_REQUIRE(head_dim == BHD, "bert_attention: head_dim %d (bert-base uses 64)", head_dim);
  const int smem = bert_attn_smem(S, false);
  MMDA_REQUIRE(B > 0 && S > 0 && smem <= 200 * 1024,
               "bert_attention: sequence %d does not fit the shared-memory resident kernel", S);
  MMDA_CUDA(cudaFuncSetAttribute(bert_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  bert_attn_fwd_kernel<<<B * nhead, ATT_THREADS, smem, stream>>>(qkv, mask, ctx, probs, S, nhead,
                                                         1.0f / sqrtf((float)head_dim), p_drop,
                                                         seed, seed_dev, stream_id);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

int mmda_bert_attention_backward(const float* qkv, const float* probs, const float* dctx,
                                 float* dqkv, int B, int S, int nhead, int head_dim, float p_drop,
                                 unsigned long long seed, const unsigned long long* seed_dev,
                                 unsigned stream_id, cudaStream_t stream) {
  MMDA_REQUIRE(head_dim == BHD, "bert_attention: head_dim %d (bert-base uses 64)", head_dim);
  const int smem = bert_attn_smem(S, true);
  MMDA_REQUIRE(B > 0 && S > 0 && smem <= 200 * 1024,
               "bert_attention: sequence %d does not fit the shared-memory resident kernel", S);
  MMDA_CUDA(cudaFuncSetAttribute(bert_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  bert_attn_bwd_kernel<<<B * nhead, ATT_THREADS, smem, stream>>>(qkv, probs, dctx, dqkv, S, nhead,
                                                         1.0f / sqrtf((float)head_dim), p_drop,
                                                         seed, seed_dev, stream_id);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

// bf16 tensor-core variants (S <= 64): same contract as the fp32 calls above; probs stays fp32.
int mmda_bert_attention_forward_mma(const float* qkv, const long long* mask, float* ctx, void* ctx_bf16,
                                    float* probs, int B, int S, int nhead, int head_dim, float p_drop,
                                    unsigned long long seed, const unsigned long long* seed_dev,
                                    unsigned stream_id, cudaStream_t stream) {
  MMDA_REQUIRE(head_dim == BHD, "bert_attention: head_dim %d (bert-base uses 64)", head_dim);
  MMDA_REQUIRE(B > 0 && S > 0 && S <= ATT_MMA_S, "bert_attention_mma: sequence %d > %d", S, ATT_MMA_S);
  MMDA_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 7) == 0,
               "bert_attention_mma: unaligned operands");
  MMDA_REQUIRE(ctx != nullptr || ctx_bf16 != nullptr, "bert_attention_mma: no output");
  bert_attn_fwd_mma_kernel<<<B * nhead, 128, 0, stream>>>(qkv, mask, ctx,
                                                          reinterpret_cast<__nv_bfloat16*>(ctx_bf16), probs, S, nhead,
                                                          1.0f / sqrtf((float)head_dim), p_drop, seed,
                                                          seed_dev, stream_id);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

int mmda_bert_attention_backward_mma(const float* qkv, const float* probs, const float* dctx,
                                     float* dqkv, void* dqkv_bf16, int B, int S, int nhead,
                                     int head_dim, float p_drop,
                                     unsigned long long seed, const unsigned long long* seed_dev,
                                     unsigned stream_id, cudaStream_t stream) {
  MMDA_REQUIRE(head_dim == BHD, "bert_attention: head_dim %d (bert-base uses 64)", head_dim);
  MMDA_REQUIRE(B > 0 && S > 0 && S <= ATT_MMA_S, "bert_attention_mma: sequence %d > %d", S, ATT_MMA_S);
  MMDA_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dctx)) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(dqkv) & 7) == 0 && (reinterpret_cast<uintptr_t>(dqkv_bf16) & 3) == 0,
               "bert_attention_mma: unaligned operands");
  MMDA_REQUIRE(dqkv != nullptr || dqkv_bf16 != nullptr, "bert_attention_mma: no output");
  MMDA_CUDA(cudaFuncSetAttribute(bert_attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 ATT_MMA_BWD_SMEM));
  bert_attn_bwd_mma_kernel<<<B * nhead, 128, ATT_MMA_BWD_SMEM, stream>>>(
      qkv, probs, dctx, dqkv, reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), S, nhead,
      1.0f / sqrtf((float)head_dim), p_drop, seed, seed_dev, stream_id);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

}  // extern "C"
