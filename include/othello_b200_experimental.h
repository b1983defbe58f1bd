/*
 * othello_b200_experimental.h -- measurement hooks of libothello_b200.so that are NOT part of
 * the drop-in boundary (othello_b200.h).  Nothing in the reference corresponds to them; bench.py
 * and tools/ use them to time the engine's own kernels inside its pipeline.  They may change
 * between ABI versions without notice.
 */
#ifndef OTHELLO_B200_EXPERIMENTAL_H
#define OTHELLO_B200_EXPERIMENTAL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-launch kernel timing, one handle per engine (no global state): create a handle, store it in
 * oth_mcts_buffers.profile, and every oth_mcts_step / oth_mcts_step_fused call on those buffers
 * records CUDA events on its stream before the step kernel, after it and after the move kernel,
 * for up to max_launches calls (do not attach a handle while the stream is being captured into a
 * graph).  oth_mcts_profile_read waits for the recorded launches and writes their durations in
 * milliseconds -- step_ms[i] = the step kernel, move_ms[i] = the move kernel of call i (either may
 * be NULL) -- and *n_launches = calls recorded; it leaves the handle empty and re-usable.
 * A handle is used by one host thread at a time (the thread that launches on the engine). */
int oth_mcts_profile_create(int32_t max_launches, void** out_handle);
int oth_mcts_profile_read(void* handle, float* step_ms, float* move_ms, int32_t* n_launches);
int oth_mcts_profile_destroy(void* handle);

#ifdef __cplusplus
}
#endif
#endif /* OTHELLO_B200_EXPERIMENTAL_H */
