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
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
  }
}
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                float* __restrict__ dx, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * expf(-0.5f * v * v);
    dx[i] = dy[i] * (cdf + v * pdf);
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
constexpr int BHD = 64;          // head dim of bert-base
constexpr int BHDP = BHD + 1;    // padded row pitch: conflict-free row-vs-row dot products

__device__ __forceinline__ float drop_scale(unsigned long long seed, unsigned stream, unsigned idx,
                                            float p, float inv_keep) {
  return (p > 0.f && rng_uniform(seed, stream, idx) < p) ? 0.f : inv_keep;
}

__global__ void __launch_bounds__(128)
bert_attn_fwd_kernel(const float* __restrict__ qkv, const long long* __restrict__ mask,
                     float* __restrict__ ctx, float* __restrict__ probs, int S, int nhead,
                     float scale, float p_drop, unsigned long long seed,
                     const unsigned long long* __restrict__ seed_dev, unsigned stream) {
  extern __shared__ float sm[];
  if (seed_dev) seed += seed_dev[0] * 0x9E3779B97F4A7C15ULL;
  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int Hd = nhead * BHD, ld = 3 * Hd, SP = S + 1;
  float* Qs = sm;
  float* Ks = Qs + S * BHDP;
  float* Vs = Ks + S * BHDP;
  float* Ps = Vs + S * BHDP;     // [S][S+1]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < S * BHD; idx += blockDim.x) {
    const int s = idx / BHD, c = idx % BHD;
    const float* r = qkv + (size_t)(b * S + s) * ld + h * BHD + c;
    Qs[s * BHDP + c] = r[0];
    Ks[s * BHDP + c] = r[Hd];
    Vs[s * BHDP + c] = r[2 * Hd];
  }
  __syncthreads();
  for (int idx = tid; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx % S;
    float acc = 0.f;
#pragma unroll 16
    for (int c = 0; c < BHD; ++c) acc = fmaf(Qs[i * BHDP + c], Ks[j * BHDP + c], acc);
    // HF adds finfo.min to masked keys: the softmax weight is exactly 0
    Ps[i * SP + j] = mask[(size_t)b * S + j] != 0 ? acc * scale : -INFINITY;
  }
  __syncthreads();
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int i = warp; i < S; i += (blockDim.x >> 5)) {
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, Ps[i * SP + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float e = expf(Ps[i * SP + j] - mx);
      Ps[i * SP + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    const unsigned base = (unsigned)(((b * nhead + h) * S + i) * S);
    for (int j = lane; j < S; j += 32) {
      const float pr = Ps[i * SP + j] * inv;
      if (probs) probs[(size_t)base + j] = pr;
      Ps[i * SP + j] = pr * drop_scale(seed, stream, base + j, p_drop, inv_keep);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < S * BHD; idx += blockDim.x) {
    const int i = idx / BHD, c = idx % BHD;
    float acc = 0.f;
    for (int j = 0; j < S; ++j) acc = fmaf(Ps[i * SP + j], Vs[j * BHDP + c], acc);
    ctx[(size_t)(b * S + i) * Hd + h * BHD + c] = acc;
  }
}

__global__ void __launch_bounds__(128)
bert_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                     const float* __restrict__ dctx, float* __restrict__ dqkv, int S, int nhead,
                     float scale, float p_drop, unsigned long long seed,
                     const unsigned long long* __restrict__ seed_dev, unsigned stream) {
  extern __shared__ float sm[];
  if (seed_dev) seed += seed_dev[0] * 0x9E3779B97F4A7C15ULL;
  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int Hd = nhead * BHD, ld = 3 * Hd, SP = S + 1;
  float* Qs = sm;
  float* Ks = Qs + S * BHDP;
  float* Vs = Ks + S * BHDP;
  float* Cs = Vs + S * BHDP;     // d(ctx)
  float* Ps = Cs + S * BHDP;     // probabilities, later dropped probabilities
  float* Ds = Ps + S * SP;       // d(dropped probs), later d(scores)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned base0 = (unsigned)((b * nhead + h) * S * S);
  for (int idx = tid; idx < S * BHD; idx += blockDim.x) {
    const int s = idx / BHD, c = idx % BHD;
    const float* r = qkv + (size_t)(b * S + s) * ld + h * BHD + c;
    Qs[s * BHDP + c] = r[0];
    Ks[s * BHDP + c] = r[Hd];
    Vs[s * BHDP + c] = r[2 * Hd];
    Cs[s * BHDP + c] = dctx[(size_t)(b * S + s) * Hd + h * BHD + c];
  }
  for (int idx = tid; idx < S * S; idx += blockDim.x)
    Ps[(idx / S) * SP + idx % S] = probs[(size_t)base0 + idx];
  __syncthreads();
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int idx = tid; idx < S * S; idx += blockDim.x) {   // d(dropped probs) -> d(probs)
    const int i = idx / S, j = idx % S;
    float acc = 0.f;
#pragma unroll 16
    for (int c = 0; c < BHD; ++c) acc = fmaf(Cs[i * BHDP + c], Vs[j * BHDP + c], acc);
    Ds[i * SP + j] = acc * drop_scale(seed, stream, base0 + idx, p_drop, inv_keep);
  }
  __syncthreads();
  for (int i = warp; i < S; i += (blockDim.x >> 5)) {      // softmax backward, then dropout of P
    float r = 0.f;
    for (int j = lane; j < S; j += 32) r += Ds[i * SP + j] * Ps[i * SP + j];
    r = warp_sum(r);
    for (int j = lane; j < S; j += 32) {
      const float pr = Ps[i * SP + j];
      Ds[i * SP + j] = pr * (Ds[i * SP + j] - r) * scale;
      Ps[i * SP + j] = pr * drop_scale(seed, stream, base0 + i * S + j, p_drop, inv_keep);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < S * BHD; idx += blockDim.x) {
    const int s = idx / BHD, c = idx % BHD;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < S; ++j) {
      dq = fmaf(Ds[s * SP + j], Ks[j * BHDP + c], dq);     // dQ[s] = sum_j dS[s][j] K[j]
      dk = fmaf(Ds[j * SP + s], Qs[j * BHDP + c], dk);     // dK[s] = sum_i dS[i][s] Q[i]
      dv = fmaf(Ps[j * SP + s], Cs[j * BHDP + c], dv);     // dV[s] = sum_i Pd[i][s] dC[i]
    }
    float* o = dqkv + (size_t)(b * S + s) * ld + h * BHD + c;
    o[0] = dq;
    o[Hd] = dk;
    o[2 * Hd] = dv;
  }
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

int mmda_gelu_forward(const float* x, float* y, long long n, cudaStream_t stream) {
  if (n <= 0) return MMDA_OK;
  gelu_fwd_kernel<<<ew_grid_b((size_t)n), 256, 0, stream>>>(x, y, (size_t)n);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

int mmda_gelu_backward(const float* dy, const float* x, float* dx, long long n, cudaStream_t stream) {
  if (n <= 0) return MMDA_OK;
  gelu_bwd_kernel<<<ew_grid_b((size_t)n), 256, 0, stream>>>(dy, x, dx, (size_t)n);
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
  return (int)(((bwd ? 4 : 3) * S * BHDP + (bwd ? 2 : 1) * S * (S + 1)) * sizeof(float));
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
  bert_attn_fwd_kernel<<<B * nhead, 128, smem, stream>>>(qkv, mask, ctx, probs, S, nhead,
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
  bert_attn_bwd_kernel<<<B * nhead, 128, smem, stream>>>(qkv, probs, dctx, dqkv, S, nhead,
                                                         1.0f / sqrtf((float)head_dim), p_drop,
                                                         seed, seed_dev, stream_id);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

}  // extern "C"
