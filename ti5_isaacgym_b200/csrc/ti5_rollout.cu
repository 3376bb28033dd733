// Rollout storage: rs:59-74 `add_transitions`, dh_ppo.py:93-103 `process_env_step`, the episode bookkeeping of
// dh_on_policy_runner.py:149-168, and rs:129-173 `mini_batch_generator`.
//
// The reference copies the whole (N, H*K) observation window into a (T,N,H*K) tensor every step and gathers rows
// of it for every mini-batch.  Consecutive windows of an env share H-1 of their H frames, so this storage keeps
// the frames once — in the logs the observation kernel writes (Ti5Buffers.frame_log) — and rebuilds a window only
// when a mini-batch asks for it: frames older than the env's last reset read as zero (t1:556-559 clears them in
// the reference's deques), which the per-(step, env) count in `valid_log` encodes.
//
// store_transition_kernel: one launch per env step instead of ~14 copy_/mul/add launches and the runner's
// per-step nonzero() + .cpu() sync.  The finished-episode list needs the ascending rank of every done env: a CTA
// with a done env counts the done flags in front of its tile (N bytes at most, L2 hits) — no inter-CTA ordering.
//
// gather_minibatch_kernel: one warp per sample.  A window is H*K contiguous floats of the env's log row range
// (two pieces when the log wraps), masked per frame; loads are 4-byte and coalesced (47-float frames are only
// 4-byte aligned), eight independent loads in flight per lane.
#include "ti5_device.cuh"
#include "ti5_host.h"

namespace ti5 {

constexpr int SB = 128;

__global__ void __launch_bounds__(SB)
store_transition_kernel(const __grid_constant__ Ti5Rollout ro, const __grid_constant__ Ti5Transition tr, int s,
                        int frame_row, float gamma) {
  __shared__ int s_warp[SB / 32];
  __shared__ int s_before;
  const int N = ro.num_envs, A = ro.num_actions, tid = threadIdx.x;
  const int e = blockIdx.x * SB + tid;
  if (e == 0) ro.frame_row[s] = frame_row;

  // (N,A) rows: flat, coalesced
  const size_t base_a = (size_t)s * N * A;
  for (int i = blockIdx.x * SB + tid; i < N * A; i += gridDim.x * SB) {
    ro.actions[base_a + i] = tr.actions[i];
    ro.mu[base_a + i] = tr.action_mean[i];
    ro.sigma[base_a + i] = tr.action_sigma[i];
  }

  bool done = false;
  float rew = 0.0f;
  if (e < N) {
    const size_t i = (size_t)s * N + e;
    const float v = tr.values[e];
    rew = tr.rewards[e];
    done = tr.dones[e] != 0;
    float stored = rew;
    if (tr.time_outs) stored = rew + gamma * (v * (tr.time_outs[e] ? 1.0f : 0.0f));   // dh_ppo.py:97-98
    ro.rewards[i] = stored;
    ro.dones[i] = done ? 1 : 0;
    ro.values[i] = v;
    ro.actions_log_prob[i] = tr.actions_log_prob[e];
  }
  if (!ro.cur_reward_sum) return;

  // runner :156-168: cur_reward_sum += rewards; cur_episode_length += 1; finished episodes appended in
  // ascending env order, then their accumulators cleared
  const int any = __syncthreads_or(done);
  const int e0 = blockIdx.x * SB;
  if (any || blockIdx.x == 0) {
    // done flags in front of this tile (CTA 0: in the whole batch, to advance the list length)
    const int upto = blockIdx.x == 0 ? N : e0;
    int c = 0;
    if ((reinterpret_cast<uintptr_t>(tr.dones) & 15) == 0) {      // 16 flags per load, all loads independent
      const uint4* d4 = reinterpret_cast<const uint4*>(tr.dones);
      const int n16 = upto >> 4;
      for (int i = tid; i < n16; i += SB) {
        const uint4 w = d4[i];
        // flags are 0 / 1 bytes (torch.bool) or any non-zero byte: count non-zero bytes
        const unsigned x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          unsigned nz = x[q] | (x[q] >> 4);
          nz |= nz >> 2;
          nz |= nz >> 1;
          c += __popc(nz & 0x01010101u);
        }
      }
      for (int i = (n16 << 4) + tid; i < upto; i += SB) c += tr.dones[i] != 0;
    } else {
      for (int i = tid; i < upto; i += SB) c += tr.dones[i] != 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) s_warp[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < SB / 32; ++w) t += s_warp[w];
      s_before = t;
    }
    __syncthreads();
  }
  const int listed = ro.n_finished[s & 1];
  if (blockIdx.x == 0 && tid == 0) ro.n_finished[(s + 1) & 1] = listed + s_before;
  const int before = blockIdx.x == 0 ? 0 : s_before;
  __syncthreads();
  const BlockRank br = block_rank(done, s_warp);
  if (e < N) {
    const float sum = ro.cur_reward_sum[e] + rew;
    const float len = ro.cur_episode_length[e] + 1.0f;
    if (done) {
      const size_t j = (size_t)listed + before + br.rank;
      ro.finished_rew[j] = sum;
      ro.finished_len[j] = len;
    }
    ro.cur_reward_sum[e] = done ? 0.0f : sum;
    ro.cur_episode_length[e] = done ? 0.0f : len;
  }
}

// window of `frames` rows of `width` floats ending at log row `r`, oldest first; rows further back than `nv` are zero.
// All U addresses are formed first and the loads issued unconditionally (a cleared frame's row is still a valid
// address): with the loads under the validity branch ptxas serialised them behind one another (ncu: every
// predicate ISETP stalled on the previous load's scoreboard).
template <int WC>
__device__ __forceinline__ void gather_window(const float* __restrict__ log_e, float* __restrict__ out, int frames,
                                              int width_rt, int L, int r, int nv, int lane) {
  const int W = WC ? WC : width_rt;
  const int total = frames * W;
  constexpr int U = 8;
  for (int i0 = 0; i0 < total; i0 += 32 * U) {
    const float* src[U];
    bool keep[U];
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = min(i0 + u * 32 + lane, total - 1);
      const int h = i / W, back = frames - 1 - h;
      int row = r - back;
      row += row < 0 ? L : 0;
      keep[u] = back < nv;
      src[u] = log_e + (size_t)row * W + (i - h * W);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(src[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 32 + lane;
      if (i < total) __stcs(out + i, keep[u] ? v[u] : 0.0f);     // streamed: the batch is consumed once by the policy
    }
  }
}

template <int KC, int PC>
__global__ void __launch_bounds__(256)
gather_minibatch_kernel(const __grid_constant__ Ti5Rollout ro, const int64_t* __restrict__ idx,
                        const int32_t* __restrict__ order, int B, const __grid_constant__ Ti5Batch out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int N = ro.num_envs, H = ro.frame_stack, CH = ro.c_frame_stack, L = ro.log_len, A = ro.num_actions;
  const int K = KC ? KC : ro.num_single_obs, P = PC ? PC : ro.priv_frame;
  // consecutive warps take consecutive entries of `order`: when the caller sorted the batch rows by (env, step),
  // the warps of a CTA rebuild overlapping windows of one env and share its frames in L1 / L2
  for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < B; j += gridDim.x * wpb) {
    const int s = order ? order[j] : j;
    const int64_t flat = idx[s];
    const int t = (int)(flat / N), e = (int)(flat - (int64_t)t * N);
    const int r = ro.frame_row[t];
    const int nv = ro.valid_log[(size_t)r * N + e];
    if (out.obs)
      gather_window<KC>(ro.frame_log + (size_t)e * L * K, out.obs + (size_t)s * H * K, H, K, L, r, nv, lane);
    if (out.critic_obs)
      gather_window<PC>(ro.priv_log + (size_t)e * L * P, out.critic_obs + (size_t)s * CH * P, CH, P, L, r, min(nv, CH), lane);
    for (int i = lane; i < A; i += 32) {
      if (out.actions) out.actions[(size_t)s * A + i] = ro.actions[(size_t)flat * A + i];
      if (out.mu) out.mu[(size_t)s * A + i] = ro.mu[(size_t)flat * A + i];
      if (out.sigma) out.sigma[(size_t)s * A + i] = ro.sigma[(size_t)flat * A + i];
    }
    if (lane == 0) {
      if (out.values) out.values[s] = ro.values[flat];
      if (out.advantages) out.advantages[s] = ro.advantages[flat];
      if (out.returns) out.returns[s] = ro.returns[flat];
      if (out.actions_log_prob) out.actions_log_prob[s] = ro.actions_log_prob[flat];
    }
  }
}

}  // namespace ti5

using namespace ti5;

static int check_rollout(const Ti5Rollout* ro) {
  TI5_CHECK_ARGS(ro != nullptr);
  TI5_CHECK_ARGS(ro->num_envs > 0 && ro->num_steps > 0 && ro->num_actions > 0);
  return TI5_OK;
}

extern "C" int ti5_store_transition(const Ti5Rollout* ro, const Ti5Transition* tr, int32_t step, int32_t frame_row,
                                    float gamma, void* stream) {
  if (int rc = check_rollout(ro)) return rc;
  TI5_CHECK_ARGS(tr != nullptr);
  TI5_CHECK_ARGS(step >= 0 && step < ro->num_steps);          // rs:60-61 "Rollout buffer overflow" is the caller's
  TI5_CHECK_ARGS(frame_row >= 0 && (ro->log_len == 0 || frame_row < ro->log_len));
  TI5_CHECK_ARGS(ro->frame_row && ro->actions && ro->mu && ro->sigma && ro->rewards && ro->dones && ro->values &&
                 ro->actions_log_prob);
  TI5_CHECK_ARGS(tr->actions && tr->action_mean && tr->action_sigma && tr->values && tr->actions_log_prob &&
                 tr->rewards && tr->dones);
  TI5_CHECK_ARGS(!ro->cur_reward_sum || (ro->cur_episode_length && ro->finished_rew && ro->finished_len && ro->n_finished));
  const int blocks = (ro->num_envs + SB - 1) / SB;
  store_transition_kernel<<<blocks, SB, 0, (cudaStream_t)stream>>>(*ro, *tr, step, frame_row, gamma);
  return ti5_check_launch("ti5_store_transition");
}

extern "C" int ti5_gather_minibatch(const Ti5Rollout* ro, const int64_t* idx, const int32_t* order, int32_t B,
                                    const Ti5Batch* out, void* stream) {
  if (int rc = check_rollout(ro)) return rc;
  TI5_CHECK_ARGS(out != nullptr && B >= 0);
  if (B == 0) return TI5_OK;
  TI5_CHECK_ARGS(idx != nullptr && ro->frame_row != nullptr);
  if (out->obs || out->critic_obs) {
    TI5_CHECK_ARGS(ro->log_len >= ro->frame_stack && ro->log_len >= ro->c_frame_stack && ro->valid_log);
    TI5_CHECK_ARGS(!out->obs || ro->frame_log);
    TI5_CHECK_ARGS(!out->critic_obs || ro->priv_log);
  }
  const int wpb = 8;
  // enough warps to fill the machine several times over, capped so each keeps a few samples
  int blocks = (B + wpb - 1) / wpb;
  const int cap = ti5_sm_count() * 8 * 4;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (ro->num_single_obs == 47 && ro->priv_frame == 73)
    gather_minibatch_kernel<47, 73><<<blocks, wpb * 32, 0, st>>>(*ro, idx, order, B, *out);
  else if (ro->num_single_obs == 47)
    gather_minibatch_kernel<47, 0><<<blocks, wpb * 32, 0, st>>>(*ro, idx, order, B, *out);
  else
    gather_minibatch_kernel<0, 0><<<blocks, wpb * 32, 0, st>>>(*ro, idx, order, B, *out);
  return ti5_check_launch("ti5_gather_minibatch");
}
