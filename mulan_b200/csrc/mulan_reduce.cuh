// Deterministic reduction of the per-example loss terms to the six loss_fn scalars
// (ldm/model_mulan_epsilon.py:357-363, ldm/experiment_vdm.py:62-74), shared by the stand-alone
// mulan_bpd_reduce kernel and the epilogue of the post kernels (mulan_post_bpd).
//
// Fixed summation order, independent of which CTA does the work and of launch timing:
//   group g = rows [128 g, 128 g + 128): lane i of a 256-thread CTA holds row 128 g + i (zero
//             for i >= group size), warp-shuffle tree, warps added in order  -> partial[g][5]
//   final   : thread i adds partial[i], partial[i + 256], ... in increasing g, then the same
//             CTA-wide tree                                                  -> 5 sums
// Parallel form ("last CTA done", CUDA threadFenceReduction pattern): the CTA that completes a
// group (per-group counter) reduces it; the CTA that completes the last group (global counter)
// does the final step.  Counters live in a caller-supplied workspace that is zero before its
// first use and is left zero by every call.  Single-CTA form (no workspace): one CTA walks the
// groups and reproduces the same order bit for bit.
#pragma once

#include "mulan_kernels.h"

namespace mulan {

constexpr int kRedGroup = 128;

__host__ __device__ inline int red_groups(int rows) { return (rows + kRedGroup - 1) / kRedGroup; }
// workspace layout (32-bit words): [0] groups done, [1 .. G] rows done per group, padding to a
// multiple of 4 words, then partial[G][8] floats
__host__ __device__ inline int red_partials_offset(int groups) { return (groups + 1 + 3) / 4 * 4; }

// Sum of one group's rows; valid in thread 0 (all 256 threads call).  Also writes
// loss_klz_total = kl_z + loss_klz_prior (ldm/model_mulan_epsilon.py:359) for the group's rows.
template <int NW = kWarps>
__device__ __forceinline__ void red_group_sum(const BpdReduceParams& p, int g, float (&acc)[5],
                                              float (*red)[5]) {
  const int i = threadIdx.x, row = g * kRedGroup + i;
#pragma unroll
  for (int k = 0; k < 5; ++k) acc[k] = 0.f;
  if (i < kRedGroup && row < p.rows) {
    // __ldcg: loss_diff may have been written by other CTAs of the running kernel
    const float klz = p.kl_z != nullptr ? __ldcg(p.kl_z + row) + __ldcg(p.loss_klz_prior + row)
                                        : __ldcg(p.loss_klz_prior + row);
    if (p.loss_klz_total != nullptr) p.loss_klz_total[row] = klz;
    acc[0] = __ldcg(p.loss_recon + row);
    acc[1] = klz;
    acc[2] = p.loss_diff != nullptr ? __ldcg(p.loss_diff + row) : 0.f;
    acc[3] = __ldcg(p.var_sums + 2 * row);
    acc[4] = __ldcg(p.var_sums + 2 * row + 1);
  }
  __syncthreads();          // `red` may still be read by thread 0 from a previous call
  block_sum<5, NW>(acc, red);      // warps beyond the fourth add zeros: same bits for any NW
}

// The pmean of the six scalars (ldm/experiment.py:347-348) as an all-gather by peer stores: the
// ONE thread that has just written this rank's scalars bumps the rank's step counter (the word
// behind the slots of its own board, so CUDA-graph replays advance it) and stores the row
// {s0 s1 s2 tag | s3 s4 s5 tag}, tag = step, into slot [step % kBoardSlots][rank] of every rank's
// board.  Each half is ONE naturally aligned 128-bit store that carries its own tag, so no
// system-scope fence sits on the kernel's tail: a reader accepts a row when both tags match.
// No collective call, no extra launch; mulan_scalar_board_read averages on demand.
__device__ __forceinline__ void board_publish(const ScalarBoard& b, const float* scalars) {
  unsigned* counter = reinterpret_cast<unsigned*>(b.boards[b.rank] + kBoardSlots * 8 * kBoardRow);
  const unsigned step = *counter + 1u;
  *counter = step;
  const int slot = (int)(step % (unsigned)kBoardSlots);
  const float tag = __uint_as_float(step);
  const float4 lo = make_float4(scalars[0], scalars[1], scalars[2], tag);
  const float4 hi = make_float4(scalars[3], scalars[4], scalars[5], tag);
  for (int r = 0; r < b.world; ++r) {
    float4* row = reinterpret_cast<float4*>(b.boards[r] + ((size_t)slot * 8 + b.rank) * kBoardRow);
    asm volatile("st.volatile.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(row), "f"(lo.x), "f"(lo.y),
                 "f"(lo.z), "f"(lo.w) : "memory");
    asm volatile("st.volatile.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(row + 1), "f"(hi.x),
                 "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
  }
}

__device__ __forceinline__ void red_write_scalars(const BpdReduceParams& p, const float (&acc)[5]) {
  const float n = (float)p.rows;
  const float rescale = (float)(1.0 / ((double)p.dim * 0.6931471805599453));
  const float bpd_recon = __fdiv_rn(acc[0], n) * rescale;
  const float bpd_latent = __fdiv_rn(acc[1], n) * rescale;
  const float bpd_diff = __fdiv_rn(acc[2], n) * rescale;
  p.scalars[0] = bpd_recon + bpd_latent + bpd_diff;
  p.scalars[1] = bpd_latent;
  p.scalars[2] = bpd_recon;
  p.scalars[3] = bpd_diff;
  const float nd = (float)((double)p.rows * (double)p.dim);
  p.scalars[4] = __fdiv_rn(acc[3], nd);
  p.scalars[5] = __fdiv_rn(acc[4], nd);
  if (p.board.world > 0) board_publish(p.board, p.scalars);
}

// Called by every thread of a CTA (NW warps, >= 8) that has just made `done_rows` more rows of
// group g final (their loss_diff is written and fenced by thread 0 before the call).
// Whole-CTA uniform.
template <int NW = kWarps>
__device__ __forceinline__ void red_rows_done(const BpdReduceParams& p, int g, int done_rows,
                                              float (*red)[5], int* s_flag) {
  const int G = red_groups(p.rows);
  float* partials = reinterpret_cast<float*>(p.ws + red_partials_offset(G));
  const int gsize = min(kRedGroup, p.rows - g * kRedGroup);
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(p.ws + 1 + g, (unsigned)done_rows);
    *s_flag = (prev + (unsigned)done_rows == (unsigned)gsize) ? 1 : 0;
  }
  __syncthreads();
  if (*s_flag == 0) return;
  __threadfence();
  float acc[5];
  red_group_sum<NW>(p, g, acc, red);
  if (G == 1) {
    // one group: its sum IS the final sum (0 + x and a tree of zeros leave x unchanged), so the
    // second round trip through the partials is skipped -- the small-batch (<= 128 rows) case
    if (threadIdx.x == 0) {
      red_write_scalars(p, acc);
      p.ws[1] = 0;
    }
    return;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) __stcg(partials + 8 * g + k, acc[k]);
    p.ws[1 + g] = 0;                                   // leave the workspace zero
    __threadfence();
    const unsigned prev = atomicAdd(p.ws, 1u);
    *s_flag = (prev + 1u == (unsigned)G) ? 2 : 0;
  }
  __syncthreads();
  if (*s_flag != 2) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < 5; ++k) acc[k] = 0.f;
  if (threadIdx.x < kThreads) {       // the canonical order is defined on 256 accumulators
    for (int gg = threadIdx.x; gg < G; gg += kThreads) {
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[k] += __ldcg(partials + 8 * gg + k);
    }
  }
  __syncthreads();
  block_sum<5, NW>(acc, red);
  if (threadIdx.x == 0) {
    red_write_scalars(p, acc);
    p.ws[0] = 0;
  }
}

}  // namespace mulan
