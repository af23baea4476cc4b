/* prelude.h -- makes an OpenCL C 1.2 kernel source compile as plain C (gcc) for oracle/minicl.
 * TEST INFRASTRUCTURE ONLY.  Nothing here is reference code: it is the compatibility layer that lets the
 * reference's *own* kernel strings (handed to clCreateProgramWithSource at run time by the reference binary)
 * execute on the CPU.  One work-group runs at a time per host thread, work-items are fibers, so `__local`
 * maps to thread-local static storage. */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define __kernel
#define __global
#define __constant const
#define __private
#define __local static __thread

typedef struct { float x, y; } float2;
typedef struct { float x, y, z, w; } float4;
typedef struct { double x, y; } double2;
typedef struct { double x, y, z, w; } double4;
static inline float2 mk_float2(float a, float b) { float2 r = {a, b}; return r; }
static inline float4 mk_float4(float a, float b, float c, float d) { float4 r = {a, b, c, d}; return r; }
static inline double2 mk_double2(double a, double b) { double2 r = {a, b}; return r; }

#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2

struct minicl_wi { size_t gid[3], lid[3], grp[3], gsz[3], lsz[3], off[3], ngrp[3]; unsigned dim; };
__thread const struct minicl_wi *minicl_cur;
void (*minicl_barrier_fn)(void);
void minicl_set_cur(const struct minicl_wi *w) { minicl_cur = w; }
void minicl_set_barrier(void (*f)(void)) { minicl_barrier_fn = f; }

static inline size_t get_global_id(unsigned d) { return minicl_cur->gid[d]; }
static inline size_t get_local_id(unsigned d) { return minicl_cur->lid[d]; }
static inline size_t get_group_id(unsigned d) { return minicl_cur->grp[d]; }
static inline size_t get_global_size(unsigned d) { return minicl_cur->gsz[d]; }
static inline size_t get_local_size(unsigned d) { return minicl_cur->lsz[d]; }
static inline size_t get_num_groups(unsigned d) { return minicl_cur->ngrp[d]; }
static inline size_t get_global_offset(unsigned d) { return minicl_cur->off[d]; }
static inline unsigned get_work_dim(void) { return minicl_cur->dim; }
static inline void barrier(int flags) { (void)flags; minicl_barrier_fn(); }
