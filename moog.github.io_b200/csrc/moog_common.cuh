// moog_common.cuh -- shared declarations of the libmoog_b200 translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/moog_b200.h"

namespace moog {

// Parsed view of a program blob (pointers may be host or device).
struct ProgramView {
  const int32_t *hdr;
  const moog_op *ops;
  const int32_t *ipool;
  const moog_ex *expr;
  const int32_t *voff;
  const double *dpool;
};

__host__ __device__ inline ProgramView view_of(const void *blob) {
  ProgramView v;
  v.hdr = (const int32_t *)blob;
  v.ops = (const moog_op *)(v.hdr + MOOG_HDR_WORDS);
  v.ipool = (const int32_t *)(v.ops + v.hdr[MOOG_H_N_OPS]);
  int npool = (v.hdr[MOOG_H_N_IPOOL] + 1) & ~1;
  v.expr = (const moog_ex *)(v.ipool + npool);
  v.voff = v.ipool + v.hdr[MOOG_H_VOFF];
  v.dpool = (const double *)(v.expr + v.hdr[MOOG_H_N_EXPR]);
  return v;
}

enum StepMode { MODE_ENV_STEP = 0, MODE_PHYSICS = 1, MODE_POST_RESET = 2, MODE_OVERLAP = 3 };

struct StepArgs {
  const void *blob;  // device copy of the program
  const uint16_t *vmap;  // [VT] cached vertex -> slot << 8 | index within the slot's outline (255: beyond 254)
  moog_state st;
  int n_envs;
  int mode;
  int S, L, K, VT, NF, CMW;  // program dimensions (filled in by launch_step from the header)
  int dcv_tile;              // contained vertices per pass of the directed search (launch_step)
  const int *order;          // [n_envs] env handled by CTA i, or nullptr = identity
  int first, count;          // this launch covers dispatch positions [first, first + count)
  int *cost;                 // [n_envs] SM cycles >> 6 this call cost each env, or nullptr
  int *done;                 // nullptr or [1 + n_envs], zeroed: count, then env + 1 in finishing order
  int *sm_active;            // nullptr or [256], zeroed: envs being stepped on each SM right now
  int trace;                 // MOOG_TRACE_TIMES=1: io.counters[n][6..7] = globaltimer at start / finish
  moog_step_io io;
  // io.frames, drawn by the step kernel itself (launch_step decides): shared-memory offset of
  // the renderer's buffers behind the part of the env record it reads
  int fused_render, render_off;
  uint8_t *frames;
  moog_state pool;  // valid iff io.pool != nullptr
  // MODE_OVERLAP
  int layer_a, layer_b;
  uint8_t *overlap_out;
};

struct RenderArgs {
  const void *blob;
  moog_state st;
  int n_envs;
  uint8_t *frames;
  long long *trace;     // MOOG_TRACE_TIMES=1: the step's counters; [n][4..5] = globaltimer at render start / end
  const int *resample;  // anti_aliasing > 1: device copy of resample_tables()
  int ksize_h, ksize_v;
};

// Launch options of one program (moog_program_set_option / the MOOG_* environment variables, which
// are read ONCE, when the program is created).  -1 / 0 = decided by the library.
struct LaunchOptions {
  int helper = -1;            // "helper": second warp per env for one direction of _get_collision_vectors
  int ctas_per_sm = 0;        // "ctas_per_sm": envs resident per SM (the shared-memory request is padded to cap it)
  int smem_pad = 0;           // "smem_pad": bytes added to the request instead
  int fused_render = -1;      // "fused_render": the step CTAs draw their env's frame themselves
  int tail_render = 2;        // "tail_render": 0 render after the step, 1 / 2 behind it (programmatic launch; 2 persistent)
  int tail_ctas_per_sm = 0;   // "tail_ctas_per_sm"
  int tail_busy_thr = 1;      // "tail_busy_thr": envs stepped on an SM from which its render CTAs stand back
  int render_epb = 0;         // "render_epb": envs per CTA of the stand-alone render kernel
  int trace_times = 0;        // "trace_times": per-env globaltimer stamps in io.counters (diagnostic)
  int seed = 0;               // "seed": the draws of moog_env_post_reset (which has no io.seed of its own)
};
LaunchOptions launch_options_from_env();
bool set_launch_option(LaunchOptions &o, const char *name, int value);

// host-side launchers (defined in the .cu files)
constexpr int kMaxForceOps = 32;
int env_smem_bytes(const int32_t *hdr, bool helper = true);
int candidate_matrix_words(const void *host_blob);
// [first, first + count): dispatch positions covered by this launch (count < 0: all);
// resident_envs_per_sm > 0 pads the shared-memory request so that at most that many envs
// share an SM; helper: every env's CTA gets a second warp that runs one direction of
// _get_collision_vectors next to its owner
struct StepPlan { size_t smem; bool fuse; int render_off; };
// frames_mode: 0 no frames, 1 frames (fused only when MOOG_FUSED_RENDER=1), 2 frames, fused preferred
StepPlan plan_step(const int32_t *host_hdr, int resident_envs_per_sm, bool helper, int frames_mode,
                   const LaunchOptions &opt = LaunchOptions());
// *fused (optional): whether the kernel also drew a.io.frames (else the caller runs launch_render)
cudaError_t launch_step(const StepArgs &a, const int32_t *host_hdr, cudaStream_t stream, int *n_launches,
                        int first = 0, int count = -1, int resident_envs_per_sm = 0, bool helper = false,
                        int frames_mode = 0, bool *fused = nullptr, const LaunchOptions &opt = LaunchOptions());
cudaError_t launch_order(const int *cost, int *order, int n, cudaStream_t stream, int *n_launches);
int resample_tables(int H, int W, int OH, int OW, std::vector<int> &table, int *ksize_h, int *ksize_v);
// done: nullptr, or the finished-env list of the step kernel launched just before on `stream`
// (StepArgs::done): the render kernel is then launched with programmatic stream serialization,
// starts while the last envs are still being stepped and draws the envs in finishing order
cudaError_t launch_render(const RenderArgs &a, const int32_t *host_hdr, cudaStream_t stream, int *n_launches,
                          int *done = nullptr, int n_done = 0, int tail_mode = 1,
                          const LaunchOptions &opt = LaunchOptions());
// ints behind the finished list that launch_render's tail kernels use: a ticket counter and the
// per-SM count of envs being stepped
constexpr int kTailExtraInts = 1 + 256;

}  // namespace moog
