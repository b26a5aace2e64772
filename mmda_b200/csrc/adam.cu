// Fused clip_grad_value_ + Adam step over one flat fp32 parameter arena (reference
// src/solver.py:185-186: clamp every gradient element to [-clip, clip], then
// torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0, amsgrad=False)).  One pass:
// 16 B read (p, g, m, v) + 12 B written (p, m, v) per parameter, float4-vectorised, HBM bound.
// Restated in oracle/explicit.py::adam_clip_step.
#include "common.cuh"

// Per-step scalars kept on the device so a captured CUDA graph can be replayed step after step:
// the step counter (also the dropout seed offset), beta^t running products and the two
// bias-correction scalars of Adam.
struct StepState {
  unsigned long long counter;
  double b1pow, b2pow;
  float step_size, inv_sqrt_bc2;
};

__global__ void step_state_advance_kernel(StepState* st, double lr, double b1, double b2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->counter += 1ULL;
    st->b1pow *= b1;
    st->b2pow *= b2;
    st->step_size = (float)(lr / (1.0 - st->b1pow));
    st->inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - st->b2pow));
  }
}

__global__ void __launch_bounds__(256)
adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, size_t n, float clip, float step_size, float b1, float b2,
                 float inv_sqrt_bc2, float eps, float grad_scale,
                 const StepState* __restrict__ state) {
  if (state != nullptr) { step_size = state->step_size; inv_sqrt_bc2 = state->inv_sqrt_bc2; }
  const size_t n4 = n >> 2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define MMDA_ADAM1(P, G, M, V)                                    \
  {                                                               \
    const float gc = fminf(fmaxf((G) * grad_scale, -clip), clip); \
    M = b1 * (M) + (1.f - b1) * gc;                               \
    V = b2 * (V) + (1.f - b2) * gc * gc;                          \
    const float denom = sqrtf(V) * inv_sqrt_bc2 + eps;            \
    P = (P) - step_size * ((M) / denom);                          \
  }
    MMDA_ADAM1(pp.x, gg.x, mm.x, vv.x)
    MMDA_ADAM1(pp.y, gg.y, mm.y, vv.y)
    MMDA_ADAM1(pp.z, gg.z, mm.z, vv.z)
    MMDA_ADAM1(pp.w, gg.w, mm.w, vv.w)
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n not a multiple of 4)
  for (size_t i = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float P = p[i], M = m[i], V = v[i];
    MMDA_ADAM1(P, g[i], M, V)
    p[i] = P; m[i] = M; v[i] = V;
  }
#undef MMDA_ADAM1
}

extern "C" int mmda_adam_clip_step(float* params, const float* grads, float* exp_avg,
                                   float* exp_avg_sq, long long n, int step, float lr, float clip,
                                   float beta1, float beta2, float eps, float grad_scale,
                                   const void* state_dev, cudaStream_t stream) {
  if (n <= 0) return MMDA_OK;
  MMDA_REQUIRE(step >= 1 || state_dev != nullptr, "adam: step must be >= 1 (got %d)", step);
  if (step < 1) step = 1;
  MMDA_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
               "adam: arena pointers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  size_t blocks = ((size_t)n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  adam_clip_kernel<<<(int)blocks, 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, (size_t)n,
                                                    clip, step_size, beta1, beta2, inv_sqrt_bc2, eps,
                                                    grad_scale, reinterpret_cast<const StepState*>(state_dev));
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

// state_dev: 32-byte device buffer, initialise with mmda_step_state_init; advance once per step
// (inside the captured graph) before the dropout sites and the optimiser read it.
extern "C" int mmda_step_state_advance(void* state_dev, float lr, float beta1, float beta2,
                                       cudaStream_t stream) {
  MMDA_REQUIRE(state_dev != nullptr, "step_state_advance: null state");
  step_state_advance_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<StepState*>(state_dev), lr, beta1, beta2);
  MMDA_CHECK_LAUNCH();
  return MMDA_OK;
}

extern "C" int mmda_step_state_init(void* state_dev, long long step, float beta1, float beta2,
                                    cudaStream_t stream) {
  StepState h;
  h.counter = (unsigned long long)step;
  h.b1pow = pow((double)beta1, (double)step);
  h.b2pow = pow((double)beta2, (double)step);
  h.step_size = 0.f;
  h.inv_sqrt_bc2 = 0.f;
  MMDA_CUDA(cudaMemcpyAsync(state_dev, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
  MMDA_CUDA(cudaStreamSynchronize(stream));
  return MMDA_OK;
}
