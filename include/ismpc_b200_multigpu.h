/*
 * ismpc_b200_multigpu.h -- the batched MPCSolver::solve on ALL GPUs of one box behind one C ABI
 * (lib/libismpc_b200_mg.so; built on include/ismpc_b200.h, the CUDA runtime and NCCL).
 *
 * the path shards trivially -- instances are independent -- so a group is
 *   * one handle + one stream + one host thread per GPU,
 *   * contiguous instance ranges [g*B/G, (g+1)*B/G) per GPU g (sizes differ by at most one, earlier GPUs take the extras),
 *   * inputs scattered once, NO per-tick communication (closed-loop state stays resident per GPU),
 *   * one collective at the end of a run: ncclAllGather of the result records (ismpc_group_formc_gather).
 * The reference has nothing of this (single process, single thread, CPU: AMR_code_DART/Controller.cpp:346-348 calls
 * solver->solve once per 10 ms tick); the group is what a host stepping many robots per tick puts where the
 * reference has its single MPCSolver object.  host/MPCSolverMultiGpu.hpp is the C++ face of it.
 *
 * Conventions as in ismpc_b200.h: plain pointers and sizes, 0 or a negative ISMPC_ERR_* code, no exit(), no CPU
 * fallback.  Host buffers should come from ismpc_host_alloc (pinned) for asynchronous copies.  A group is not
 * thread-safe: one caller thread at a time (its own worker threads are internal).
 */
#ifndef ISMPC_B200_MULTIGPU_H
#define ISMPC_B200_MULTIGPU_H

#include "ismpc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ismpc_group ismpc_group;

enum { ISMPC_GATHER_NCCL = 0, ISMPC_GATHER_HOST = 1 };

/* devices: CUDA device ordinals, one shard each.  gather_mode: ISMPC_GATHER_NCCL (a communicator over the devices,
 * ncclCommInitAll; needs distinct devices) or ISMPC_GATHER_HOST (per-shard copies through the host: for a single GPU, or
 * several shards on one device in tests).  max_batch_per_device bounds every shard. */
int ismpc_group_create(ismpc_group** out, const int* devices, int n_devices, int max_batch_per_device, int gather_mode);
int ismpc_group_destroy(ismpc_group* g);
int ismpc_group_size(const ismpc_group* g);
const char* ismpc_group_last_error(const ismpc_group* g);
int64_t ismpc_group_kernel_launches(const ismpc_group* g);
/* The per-device handle of shard `rank` (e.g. for ismpc_set_option); NULL if out of range. */
ismpc_handle* ismpc_group_handle(ismpc_group* g, int rank);
/* Contiguous shard of `rank` for n_total instances: [*first, *first + *count). */
int ismpc_group_shard(const ismpc_group* g, int n_total, int rank, int* first, int* count);

/* MPCSolver::MPCSolver on every device: model tables, per-mpcIter gait tables, the footstep plans (replicated: they are
 * constructor data, MPCSolver.cpp:5, and read-only). */
int ismpc_group_formc_configure(ismpc_group* g, const ismpc_formc_model_t* model, int S, int F_ds,
                                const double* plan_xyzt, int plan_rows);

/* One tick of MPCSolver::solve for n_total instances with HOST buffers: every device takes its shard (copy in, kernel,
 * copy out on its own stream, driven by its own host thread); returns when every shard's records are in `out`. */
int ismpc_group_formc_solve_batch(ismpc_group* g, int n_total, const ismpc_state_t* state, const ismpc_walk_t* walk,
                                  const ismpc_formc_inst_t* inst, ismpc_formc_out_t* out);

/* The same with packed tick records (ismpc_formc_solve_batch_packed): ismpc_group_formc_set_instances hands every device
 * the constants of its shard of the n_total instances once; a tick then moves one 128-byte {State, WalkState} record per
 * instance in and one result record out -- read and written in place by each device's kernel when `tick` / `out` are
 * pinned (ismpc_host_alloc) and 128-byte aligned, one kernel launch per device and tick. */
int ismpc_group_formc_set_instances(ismpc_group* g, int n_total, const ismpc_formc_inst_t* inst);
int ismpc_group_formc_solve_batch_packed(ismpc_group* g, int n_total, const ismpc_formc_tick_t* tick, ismpc_formc_out_t* out);

/* Closed loop with the state resident per GPU: scatter once, advance with no communication, gather once.
 * push (nullable): n_total entries, ticks counted from the scatter. */
int ismpc_group_formc_scatter(ismpc_group* g, int n_total, const ismpc_state_t* state, const ismpc_walk_t* walk,
                              const ismpc_formc_inst_t* inst, const ismpc_push_t* push);
/* n_ticks closed-loop ticks of every resident shard (ismpc_formc_rollout per device); per-instance status bits are OR-ed
 * into the resident status words.  Asynchronous: returns once every device has its work enqueued; ismpc_group_wait or
 * ismpc_group_formc_gather wait for it.  Pushes are applied by tick count since the scatter only within the first call. */
int ismpc_group_formc_rollout(ismpc_group* g, int n_ticks);
int ismpc_group_wait(ismpc_group* g);
/* The one collective of the path: all-gather of the resident (state, walk, status) records over the devices
 * (ncclAllGather on every device's stream; every device ends up with all n_total records), then one copy from the first
 * device to the host arrays (each nullable). */
int ismpc_group_formc_gather(ismpc_group* g, ismpc_state_t* state_out, ismpc_walk_t* walk_out, int32_t* status_out);

#ifdef __cplusplus
}
#endif
#endif /* ISMPC_B200_MULTIGPU_H */
