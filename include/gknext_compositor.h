/* gknext_compositor.h — the multi-GPU frame compositor of SURVEY.md §8(e) as a C ABI (lib/libgknext_comp.so).
 *
 * The reference is single-GPU: there is no reference call these replace.  They are the frame-end step a multi-process host
 * (one process per GPU, scene and BVH replicated, image rows partitioned in interleaved tiles) puts between
 * gk_trace_frame and gk_filter_frame of include/gknext_cuda.h:
 *
 *   rank 0:      gkc_get_unique_id(id)            -> the launcher hands `id` to every rank (MPI, a file, torch.distributed ...)
 *   every rank:  gkc_create(ctx, rank, world, id) -> NCCL communicator on the context's device; the CUDA IPC handles of the
 *                                                    exchange planes are all-gathered over NCCL and the peers' planes mapped
 *   per frame:   gk_trace_frame(ctx); gkc_composite_frame(comp); gk_filter_frame(ctx);               (temporal + JBF frames)
 *            or  gk_trace_frame(ctx); gk_filter_frame_owned(ctx); gkc_composite_final(comp, dst);    (progressive, no denoiser)
 *            or  gk_trace_frame(ctx) of frame f0+rank; gkc_composite_frame_shard(comp, dst);          (frame sharding)
 *   every rank:  gkc_destroy(comp)                -> collective: unmap the peers, barrier, free
 *
 * Everything is ordered on gk_stream(ctx): a barrier is a 4-byte ncclAllReduce on that stream, the row exchange is the
 * library's peer-to-peer store kernel over NVLink (gk_exchange_push / _push_final / gk_frame_shard_push).  No host
 * synchronisation happens inside a composite call.  All functions return 0 or a negative status with gkc_last_error().
 * Collective calls must be made by every rank in the same order.  One compositor per context; not thread-safe. */
#ifndef GKNEXT_COMPOSITOR_H
#define GKNEXT_COMPOSITOR_H
#include "gknext_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif

#define GKC_UNIQUE_ID_BYTES 128
typedef struct GkCompositor GkCompositor;

const char* gkc_last_error(void);
int gkc_get_unique_id(void* id, size_t bytes);
int gkc_create(GkContext* ctx, int rank, int world, const void* id, size_t bytes, GkCompositor** out);
/* maps the gather buffers of frame-sharded rendering (contexts created with GK_CFG_TRACE_ALL_ROWS); collective */
int gkc_enable_frame_sharding(GkCompositor* comp);
int gkc_barrier(GkCompositor* comp);
/* barrier -> gk_exchange_push (six integrator planes, 44 B/px of the owned rows to every peer) -> barrier */
int gkc_composite_frame(GkCompositor* comp);
/* barrier -> gk_exchange_push_final (8 B/px of the owned rows of rtDenoised to dst_rank, -1 = everybody) -> barrier */
int gkc_composite_final(GkCompositor* comp, int dst_rank);
/* barrier -> gk_frame_shard_push -> barrier -> gk_frame_shard_accumulate -> gk_exchange_push_final -> barrier */
int gkc_composite_frame_shard(GkCompositor* comp, int dst_rank);
/* world size / rank the compositor was created with */
int gkc_world(const GkCompositor* comp);
int gkc_rank(const GkCompositor* comp);
void gkc_destroy(GkCompositor* comp);

#ifdef __cplusplus
}
#endif
#endif
