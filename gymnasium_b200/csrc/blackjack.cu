// blackjack.cu -- fused Blackjack-v1 step + TimeLimit + autoreset kernels (sm_100a).
//
// Replaces, for a batch of n envs in one launch:
//   draw_card / draw_hand / sum_hand / is_bust / score / is_natural   gymnasium/envs/toy_text/blackjack.py:15-45
//   BlackjackEnv.step    blackjack.py:178-208   (hit: bust -> -1; stick: dealer draws to >= 17, cmp of the scores,
//                                                sab auto-win on a natural, `natural` 1.5 pay-out)
//   BlackjackEnv.reset   blackjack.py:215-238   (dealer hand, player hand, then the draws that pick the rendered suit /
//                                                face of the dealer's top card -- they advance the stream, so they are made)
//   BlackjackEnv._get_obs blackjack.py:210-213  -> (player sum, dealer's first card, usable ace)
//   TimeLimit / SyncVectorEnv autoreset as in cartpole.cu
// Card draws: np_random.choice(deck) = deck[Lemire-bounded uint32 of 13] on PCG64's buffered 32-bit words (low half of a
// 64-bit draw first, high half on the next call; oracle/np_rng.py: next_uint32 / bounded_uint32).  The one-word buffer is
// per-env state (`u32buf`: bit 32 = valid, low 32 bits = the word).  A hand is kept as (raw sum, has ace, card count):
// every rule of the reference is a function of those.  Integer state, bit-exact with the reference.
//
// Memory: HBM-bound integer work, 1 thread/env.  Per env-step: PCG64 state 16 r/w + inc 16 r, buffer 8 r/w, hand 4 r/w,
// ctrl 4 r/w, action 8, obs 24, reward 8, flags 2 = 118 B.
#include "common.cuh"

namespace b2e {
namespace {

struct JackArgs {
  int64_t n, env_offset;
  int32_t max_steps, mode, rng_mode, natural, sab;
  uint64_t philox_seed, call_counter;
  int32_t* __restrict__ hand;     // [n] packed hands, see pack()
  int64_t* __restrict__ u32buf;   // [n] PCG64 32-bit buffer
  int32_t* __restrict__ ctrl;
  uint64_t* __restrict__ rng;
  int64_t* __restrict__ obs;      // [3][n]: player sum, dealer's first card, usable ace
  double* __restrict__ reward;
  uint8_t* __restrict__ term;
  uint8_t* __restrict__ trunc;
  int64_t* __restrict__ final_obs;  // [3][n]
  const void* __restrict__ actions;
  const uint8_t* __restrict__ mask;
};

struct Hand {
  int sum, ace, cnt;  // raw sum (aces count 1), 1 if the hand holds an ace, number of cards (saturates at 3)
  __device__ __forceinline__ void add(int card) {
    sum += card;
    ace |= card == 1;
    cnt = min(cnt + 1, 3);
  }
  __device__ __forceinline__ int usable_ace() const { return ace && sum + 10 <= 21; }                 // blackjack.py:26-27
  __device__ __forceinline__ int total() const { return usable_ace() ? sum + 10 : sum; }              // :30-33
  __device__ __forceinline__ int score() const { return total() > 21 ? 0 : total(); }                 // :36-41
  __device__ __forceinline__ bool natural() const { return cnt == 2 && ace && sum == 11; }            // :44-45
};

struct Table {
  Hand player, dealer;
  int dealer_first;
};
// hand word: bits 0-5 player sum, 6 player ace, 7-8 player count, 9-14 dealer sum, 15 dealer ace, 16-17 dealer count,
// 18-21 dealer's first card
__device__ __forceinline__ int32_t pack(const Table& t) {
  return t.player.sum | t.player.ace << 6 | t.player.cnt << 7 | t.dealer.sum << 9 | t.dealer.ace << 15 | t.dealer.cnt << 16 |
         t.dealer_first << 18;
}
__device__ __forceinline__ Table unpack(int32_t w) {
  Table t;
  t.player = Hand{w & 63, (w >> 6) & 1, (w >> 7) & 3};
  t.dealer = Hand{(w >> 9) & 63, (w >> 15) & 1, (w >> 16) & 3};
  t.dealer_first = (w >> 18) & 15;
  return t;
}

// the draw source of one env for one call: numpy-parity PCG64 + 32-bit buffer, or stateless Philox words
struct Cards {
  bool numpy;
  Pcg64 g;
  bool has32;
  uint32_t word;
  uint64_t seed, env, counter;
  uint32_t k;
  uint4 block;
  __device__ __forceinline__ uint32_t next32() {
    if (numpy) {
      if (has32) {
        has32 = false;
        return word;
      }
      const uint64_t x = g.next_u64();
      has32 = true;
      word = (uint32_t)(x >> 32);
      return (uint32_t)x;
    }
    if ((k & 3u) == 0u) block = philox_block(seed, env, counter, 16u + (k >> 2));
    const uint32_t r = (k & 3u) == 0u ? block.x : (k & 3u) == 1u ? block.y : (k & 3u) == 2u ? block.z : block.w;
    ++k;
    return r;
  }
  // uniform in [0, n): Lemire's multiply-and-reject (numpy: buffered_bounded_lemire_uint32)
  __device__ __forceinline__ uint32_t below(uint32_t n) {
    uint64_t m = (uint64_t)next32() * n;
    uint32_t leftover = (uint32_t)m;
    if (leftover < n) {
      const uint32_t threshold = (0xffffffffu - (n - 1u)) % n;
      while (leftover < threshold) {
        m = (uint64_t)next32() * n;
        leftover = (uint32_t)m;
      }
    }
    return (uint32_t)(m >> 32);
  }
  __device__ __forceinline__ int card() {  // deck = [1..10, 10, 10, 10] (blackjack.py:15)
    const int idx = (int)below(13u);
    return idx < 9 ? idx + 1 : 10;
  }
};

__device__ __forceinline__ Cards open_cards(const JackArgs& a, int64_t i) {
  Cards c;
  c.numpy = a.rng_mode == B2E_RNG_NUMPY;
  c.has32 = false;
  c.word = 0;
  if (c.numpy) {
    c.g = pcg64_load(a.rng, a.n, i);
    const int64_t b = a.u32buf[i];
    c.has32 = (b >> 32) & 1;
    c.word = (uint32_t)b;
  }
  c.seed = a.philox_seed; c.env = (uint64_t)(a.env_offset + i); c.counter = a.call_counter; c.k = 0;
  return c;
}
__device__ __forceinline__ void close_cards(const JackArgs& a, int64_t i, const Cards& c) {
  if (!c.numpy) return;
  pcg64_store_state(a.rng, i, c.g);
  a.u32buf[i] = (int64_t)c.word | ((int64_t)c.has32 << 32);
}

__device__ __forceinline__ Table deal(Cards& c) {  // BlackjackEnv.reset
  Table t;
  t.dealer = Hand{0, 0, 0};
  t.player = Hand{0, 0, 0};
  const int d0 = c.card();
  t.dealer.add(d0);
  t.dealer.add(c.card());
  t.player.add(c.card());
  t.player.add(c.card());
  t.dealer_first = d0;
  (void)c.below(4u);               // dealer_top_card_suit (:227)
  if (d0 == 10) (void)c.below(3u);  // dealer_top_card_value_str in J/Q/K (:231-232)
  return t;
}
__device__ __forceinline__ void write_obs(int64_t* __restrict__ obs, int64_t n, int64_t i, const Table& t) {
  __stcs(obs + i, (int64_t)t.player.total());
  __stcs(obs + n + i, (int64_t)t.dealer_first);
  __stcs(obs + 2 * n + i, (int64_t)t.player.usable_ace());
}

__global__ void __launch_bounds__(kBlock) blackjack_reset_kernel(const JackArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  if (a.mask != nullptr && a.mask[i] == 0) return;
  Cards c = open_cards(a, i);
  const Table t = deal(c);
  close_cards(a, i, c);
  a.hand[i] = pack(t);
  a.ctrl[i] = 0;
  write_obs(a.obs, a.n, i, t);
}

template <typename ActT>
__global__ void __launch_bounds__(kBlock) blackjack_step_kernel(const JackArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int32_t c = a.ctrl[i];
  const int32_t hw = a.hand[i];
  const int action = load_action<ActT>(a.actions, i);
  Cards cards = open_cards(a, i);
  if (a.mode == B2E_AUTORESET_NEXT_STEP && ctrl_pending(c)) {  // sync_vector_env.py:279-284
    const Table t = deal(cards);
    close_cards(a, i, cards);
    a.hand[i] = pack(t);
    a.ctrl[i] = 0;
    write_obs(a.obs, a.n, i, t);
    __stcs(a.reward + i, 0.0);
    a.term[i] = 0;
    a.trunc[i] = 0;
    return;
  }
  Table t = unpack(hw);
  bool terminated;
  double reward;
  if (action != 0) {  // hit (blackjack.py:180-187); the reference asserts action in {0, 1}, other values count as hit
    t.player.add(cards.card());
    terminated = t.player.total() > 21;
    reward = terminated ? -1.0 : 0.0;
  } else {            // stick (:188-204)
    terminated = true;
    while (t.dealer.total() < 17) t.dealer.add(cards.card());
    const int ps = t.player.score(), ds = t.dealer.score();
    reward = (double)(ps > ds) - (double)(ps < ds);
    if (a.sab && t.player.natural() && !t.dealer.natural()) reward = 1.0;
    else if (!a.sab && a.natural && t.player.natural() && reward == 1.0) reward = 1.5;
  }
  const int32_t elapsed = ctrl_elapsed(c) + 1;
  const bool trunc = a.max_steps > 0 && elapsed >= a.max_steps;
  __stcs(a.reward + i, reward);
  a.term[i] = terminated;
  a.trunc[i] = trunc;
  int32_t cn = elapsed;
  if (terminated || trunc) {
    if (a.mode == B2E_AUTORESET_NEXT_STEP) {
      cn |= kPending;
    } else if (a.mode == B2E_AUTORESET_SAME_STEP) {  // sync_vector_env.py:302-319
      write_obs(a.final_obs, a.n, i, t);
      t = deal(cards);
      cn = 0;
    }
  }
  close_cards(a, i, cards);
  a.hand[i] = pack(t);
  a.ctrl[i] = cn;
  write_obs(a.obs, a.n, i, t);
}

int fill(const b2e_batch* b, const b2e_blackjack_cfg* cfg, JackArgs& a, const char* fn) {
  if (int e = check_batch(b, fn)) return e;
  if (!cfg || !cfg->hand || !cfg->ctrl || (b->rng_mode == B2E_RNG_NUMPY && (!cfg->rng || !cfg->u32buf))) {
    set_error("%s: null pointer in cfg", fn);
    return B2E_EINVAL;
  }
  a = JackArgs{};
  a.n = b->n; a.env_offset = b->env_offset; a.max_steps = b->max_episode_steps; a.mode = b->autoreset_mode;
  a.rng_mode = b->rng_mode; a.philox_seed = b->philox_seed; a.call_counter = b->call_counter;
  a.natural = cfg->natural != 0; a.sab = cfg->sab != 0;
  a.hand = cfg->hand; a.u32buf = cfg->u32buf; a.ctrl = cfg->ctrl; a.rng = cfg->rng;
  return 0;
}

}  // namespace
}  // namespace b2e

using namespace b2e;

extern "C" int b2e_blackjack_reset(const b2e_batch* b, const b2e_blackjack_cfg* cfg, const uint8_t* mask, int64_t* obs,
                                   void* stream) {
  JackArgs a;
  if (int e = fill(b, cfg, a, "b2e_blackjack_reset")) return e;
  if (!obs) {
    set_error("b2e_blackjack_reset: obs is NULL");
    return B2E_EINVAL;
  }
  if (b->n == 0) return 0;
  a.mask = mask; a.obs = obs;
  blackjack_reset_kernel<<<grid_for(b->n), kBlock, 0, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "b2e_blackjack_reset");
}

extern "C" int b2e_blackjack_step(const b2e_batch* b, const b2e_blackjack_cfg* cfg, const void* actions, int64_t* obs,
                                  double* reward, uint8_t* terminated, uint8_t* truncated, int64_t* final_obs, void* stream) {
  JackArgs a;
  if (int e = fill(b, cfg, a, "b2e_blackjack_step")) return e;
  if (!actions || !obs || !reward || !terminated || !truncated ||
      (b->autoreset_mode == B2E_AUTORESET_SAME_STEP && !final_obs)) {
    set_error("b2e_blackjack_step: null pointer");
    return B2E_EINVAL;
  }
  if (b->n == 0) return 0;
  a.actions = actions; a.obs = obs; a.reward = reward; a.term = terminated; a.trunc = truncated; a.final_obs = final_obs;
  cudaStream_t s = (cudaStream_t)stream;
  switch (b->action_dtype) {
    case B2E_ACT_I64: blackjack_step_kernel<int64_t><<<grid_for(b->n), kBlock, 0, s>>>(a); break;
    case B2E_ACT_I32: blackjack_step_kernel<int32_t><<<grid_for(b->n), kBlock, 0, s>>>(a); break;
    case B2E_ACT_U8: blackjack_step_kernel<uint8_t><<<grid_for(b->n), kBlock, 0, s>>>(a); break;
    default: set_error("b2e_blackjack_step: action_dtype %d is not a discrete dtype", b->action_dtype); return B2E_EINVAL;
  }
  return cuda_status(cudaGetLastError(), "b2e_blackjack_step");
}
