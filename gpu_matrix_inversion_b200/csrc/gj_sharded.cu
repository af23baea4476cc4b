// Column-sharded single inversion: per-rank primitives (one process per GPU).  The exchange step
// (panel broadcast) is owned by the host through torch.distributed / NCCL; see DESIGN.md.
// Round-1 status: entry points declared and exported; implementation lands after the single-GPU
// path is parity-green (they return MATINV_E_UNSUPPORTED until then).
#include "../../include/matinv_shim.h"

extern "C" {
long long matinv_shard_panel_bytes(int n) { (void)n; return 0; }
int matinv_shard_create(int n, int rank, int world, matinv_shard_t **out) { (void)n; (void)rank; (void)world; if (out) *out = nullptr; return MATINV_E_UNSUPPORTED; }
void matinv_shard_destroy(matinv_shard_t *s) { (void)s; }
float *matinv_shard_local(matinv_shard_t *s, long long *local_cols, long long *local_ld) { (void)s; (void)local_cols; (void)local_ld; return nullptr; }
int matinv_shard_generate(matinv_shard_t *s, unsigned long long seed, int kind, void *stream) { (void)s; (void)seed; (void)kind; (void)stream; return MATINV_E_UNSUPPORTED; }
int matinv_shard_factor(matinv_shard_t *s, int J, void *panel_dev, void *stream) { (void)s; (void)J; (void)panel_dev; (void)stream; return MATINV_E_UNSUPPORTED; }
int matinv_shard_apply(matinv_shard_t *s, int J, const void *panel_dev, void *stream) { (void)s; (void)J; (void)panel_dev; (void)stream; return MATINV_E_UNSUPPORTED; }
int matinv_shard_status(matinv_shard_t *s, int *info_host, int *piv_host, void *stream) { (void)s; (void)info_host; (void)piv_host; (void)stream; return MATINV_E_UNSUPPORTED; }
}
