// taxi.cu -- the fickle passenger of Taxi-v4 (gymnasium/envs/toy_text/taxi.py:436-452, :466-468) as fix-up kernels around the
// generic tabular step (frozenlake.cu), so that the hot tabular kernels stay as they are:
//   reset  : after the initial-state draw, ``fickle_step = np_random.random() < fickle_probability``          (:466-468)
//   step   : if fickle_step and the passenger was on board (state BEFORE the step) and the taxi moved, then
//            fickle_step = False; dest = np_random.choice([d for d in 0..3 if d != old dest]); s = encode(...)  (:441-452)
// Both continue the env's PCG64 stream right after the draw the tabular kernel made in the same call, which is the
// reference's order.  ``choice`` of 3 = Lemire-bounded draw on PCG64's buffered 32-bit words (blackjack.cu, oracle/np_rng.py).
// Envs that the tabular kernel reset in this call (NEXT_STEP autoreset: their ctrl word is exactly 0 afterwards, a stepped
// env has elapsed >= 1) take the reset rule.  numpy-parity RNG mode only.
#include "common.cuh"

namespace b2e {
namespace {

struct FickleArgs {
  int64_t n;
  double probability;
  const int32_t* __restrict__ prev_state;  // [n] states before the tabular step (step only)
  int32_t* __restrict__ pstate;            // [n]
  const int32_t* __restrict__ ctrl;        // [n]
  uint64_t* __restrict__ rng;
  int64_t* __restrict__ u32buf;            // [n] PCG64 32-bit word buffer
  uint8_t* __restrict__ fickle;            // [n] fickle_step flags
  int64_t* __restrict__ obs;               // [n] (step only)
  const uint8_t* __restrict__ mask;        // reset only
};

__device__ __forceinline__ void draw_fickle(const FickleArgs& a, int64_t i) {
  Pcg64 g = pcg64_load(a.rng, a.n, i);
  a.fickle[i] = g.next_double() < a.probability;
  pcg64_store_state(a.rng, i, g);
}

__global__ void __launch_bounds__(kBlock) taxi_fickle_reset_kernel(const FickleArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n || (a.mask != nullptr && a.mask[i] == 0)) return;
  draw_fickle(a, i);
}

__global__ void __launch_bounds__(kBlock) taxi_fickle_step_kernel(const FickleArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  if (a.ctrl[i] == 0) {  // the tabular kernel reset this env in this call
    draw_fickle(a, i);
    return;
  }
  if (!a.fickle[i]) return;
  // state = ((row * 5 + col) * 5 + passenger) * 4 + destination   (taxi.py:373-382)
  const int s0 = a.prev_state[i], s1 = a.pstate[i];
  const int dest0 = s0 & 3, pass0 = (s0 >> 2) % 5, cell0 = (s0 >> 2) / 5, cell1 = (s1 >> 2) / 5;
  if (pass0 != 4 || cell0 == cell1) return;
  a.fickle[i] = 0;
  // np_random.choice of the three other destinations
  Pcg64 g = pcg64_load(a.rng, a.n, i);
  int64_t b = a.u32buf[i];
  bool has32 = (b >> 32) & 1;
  uint32_t word = (uint32_t)b;
  auto next32 = [&]() {
    if (has32) {
      has32 = false;
      return word;
    }
    const uint64_t x = g.next_u64();
    has32 = true;
    word = (uint32_t)(x >> 32);
    return (uint32_t)x;
  };
  uint64_t m = (uint64_t)next32() * 3u;
  uint32_t leftover = (uint32_t)m;
  if (leftover < 3u) {
    const uint32_t threshold = (0xffffffffu - 2u) % 3u;
    while (leftover < threshold) {
      m = (uint64_t)next32() * 3u;
      leftover = (uint32_t)m;
    }
  }
  const int k = (int)(m >> 32);
  const int dest = k < dest0 ? k : k + 1;  // k-th element of [d for d in range(4) if d != dest0]
  pcg64_store_state(a.rng, i, g);
  a.u32buf[i] = (int64_t)word | ((int64_t)has32 << 32);
  const int s = (s1 & ~3) | dest;
  a.pstate[i] = s;
  a.obs[i] = s;
}

}  // namespace
}  // namespace b2e

using namespace b2e;

extern "C" int b2e_taxi_fickle_reset(const b2e_batch* b, double fickle_probability, const uint8_t* mask, uint64_t* rng,
                                     uint8_t* fickle, void* stream) {
  if (int e = check_batch(b, "b2e_taxi_fickle_reset")) return e;
  if (b->rng_mode != B2E_RNG_NUMPY || !rng || !fickle) {
    set_error("b2e_taxi_fickle_reset: needs the numpy-parity RNG mode and non-null rng / fickle");
    return B2E_EINVAL;
  }
  if (b->n == 0) return 0;
  FickleArgs a{};
  a.n = b->n; a.probability = fickle_probability; a.rng = rng; a.fickle = fickle; a.mask = mask;
  taxi_fickle_reset_kernel<<<grid_for(b->n), kBlock, 0, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "b2e_taxi_fickle_reset");
}

extern "C" int b2e_taxi_fickle_step(const b2e_batch* b, double fickle_probability, const int32_t* prev_state, int32_t* pstate,
                                    const int32_t* ctrl, uint64_t* rng, int64_t* u32buf, uint8_t* fickle, int64_t* obs,
                                    void* stream) {
  if (int e = check_batch(b, "b2e_taxi_fickle_step")) return e;
  if (b->rng_mode != B2E_RNG_NUMPY || b->autoreset_mode == B2E_AUTORESET_SAME_STEP || !prev_state || !pstate || !ctrl || !rng ||
      !u32buf || !fickle || !obs) {
    set_error("b2e_taxi_fickle_step: needs the numpy-parity RNG mode, NEXT_STEP or DISABLED autoreset and non-null buffers");
    return B2E_EINVAL;
  }
  if (b->n == 0) return 0;
  FickleArgs a{};
  a.n = b->n; a.probability = fickle_probability; a.prev_state = prev_state; a.pstate = pstate; a.ctrl = ctrl; a.rng = rng;
  a.u32buf = u32buf; a.fickle = fickle; a.obs = obs;
  taxi_fickle_step_kernel<<<grid_for(b->n), kBlock, 0, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "b2e_taxi_fickle_step");
}
