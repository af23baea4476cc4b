// minicl -- a minimal OpenCL 1.2 *CPU* runtime, just large enough to execute the reference's own host code
// and kernel strings in this GPU-less, ICD-less container (the reference links but cannot run
// because there is no OpenCL platform).
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  It lets tests/test_oracle_vs_reference.py pin the oracle
// against outputs of the UNMODIFIED reference (compiled from /root/reference into oracle/_ref/ by
// oracle/Makefile.ref).  Nothing of the product links or loads this file.
//
// How it works: clCreateProgramWithSource keeps the OpenCL C text the reference passes at run time;
// clBuildProgram prepends oracle/minicl/prelude.h, rewrites the two OpenCL-only syntaxes the kernels use
// (vector literals `(float2)(a, b)`), appends one argument-unpacking entry per `__kernel`, compiles the result
// with gcc into a shared object and dlopens it; clEnqueueNDRangeKernel runs work-groups on the host: plain
// loops for barrier-free kernels, ucontext fibers (one per work-item) for kernels that call barrier().
// The queue is in-order and synchronous.  -ffp-contract is chosen by MINICL_FP_CONTRACT (off|fast): OpenCL C
// allows either for `a - b*c`.
#define CL_TARGET_OPENCL_VERSION 120
#define CL_USE_DEPRECATED_OPENCL_1_1_APIS
#define CL_USE_DEPRECATED_OPENCL_1_2_APIS
#include <CL/cl.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <ucontext.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <regex>
#include <sstream>
#include <string>
#include <vector>

#ifndef MINICL_PRELUDE
#error "compile with -DMINICL_PRELUDE=\"/abs/path/to/prelude.h\""
#endif

struct WI { size_t gid[3], lid[3], grp[3], gsz[3], lsz[3], off[3], ngrp[3]; unsigned dim; };
typedef void (*entry_fn)(void **);
typedef void (*setcur_fn)(const WI *);
typedef void (*setbar_fn)(void (*)(void));

struct _cl_platform_id { int id; };
struct _cl_device_id { int id; };
struct _cl_context { int refs; };
struct _cl_command_queue { int refs; };
struct _cl_mem { void *host; size_t size; int refs; };
struct _cl_event { int refs; cl_ulong t0, t1; };
struct KInfo { std::string name; std::vector<int> kinds; /* 0 pointer, 4 32-bit scalar, 8 64-bit scalar */ };
struct _cl_program {
    int refs; std::string src, log; void *dl; bool built, has_barrier; std::vector<KInfo> kernels;
    setcur_fn setcur; setbar_fn setbar;
};
struct _cl_kernel {
    int refs; _cl_program *prog; KInfo info; entry_fn fn;
    std::vector<std::vector<unsigned char>> argbytes; std::vector<void *> argptr;
};

static _cl_platform_id g_platform = {1};
static _cl_device_id g_device = {1};
static const size_t MAX_WG = 1024;

static cl_int put(const void *src, size_t n, size_t cap, void *dst, size_t *ret) {
    if (ret) *ret = n;
    if (dst) { if (cap < n) return CL_INVALID_VALUE; memcpy(dst, src, n); }
    return CL_SUCCESS;
}
static cl_int put_str(const char *s, size_t cap, void *dst, size_t *ret) { return put(s, strlen(s) + 1, cap, dst, ret); }

extern "C" {

cl_int clGetPlatformIDs(cl_uint n, cl_platform_id *p, cl_uint *np) {
    if (np) *np = 1;
    if (p && n >= 1) p[0] = &g_platform;
    return CL_SUCCESS;
}
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info name, size_t cap, void *v, size_t *ret) {
    switch (name) {
        case CL_PLATFORM_VERSION: return put_str("OpenCL 1.2 minicl", cap, v, ret);
        case CL_PLATFORM_NAME: return put_str("minicl (CPU test runtime)", cap, v, ret);
        case CL_PLATFORM_VENDOR: return put_str("gpu_matrix_inversion_b200/oracle", cap, v, ret);
        case CL_PLATFORM_PROFILE: return put_str("FULL_PROFILE", cap, v, ret);
        default: return put_str("", cap, v, ret);
    }
}
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint n, cl_device_id *d, cl_uint *nd) {
    if (nd) *nd = 1;
    if (d && n >= 1) d[0] = &g_device;
    return CL_SUCCESS;
}
cl_int clGetDeviceInfo(cl_device_id, cl_device_info name, size_t cap, void *v, size_t *ret) {
    switch (name) {
        case CL_DEVICE_PLATFORM: { cl_platform_id p = &g_platform; return put(&p, sizeof(p), cap, v, ret); }
        case CL_DEVICE_TYPE: { cl_device_type t = CL_DEVICE_TYPE_GPU; return put(&t, sizeof(t), cap, v, ret); }
        case CL_DEVICE_NAME: return put_str("minicl host CPU", cap, v, ret);
        case CL_DEVICE_VENDOR: return put_str("minicl", cap, v, ret);
        case CL_DEVICE_VERSION: return put_str("OpenCL 1.2 minicl", cap, v, ret);
        case CL_DRIVER_VERSION: return put_str("1.0", cap, v, ret);
        case CL_DEVICE_OPENCL_C_VERSION: return put_str("OpenCL C 1.2", cap, v, ret);
        case CL_DEVICE_PROFILE: return put_str("FULL_PROFILE", cap, v, ret);
        case CL_DEVICE_EXTENSIONS: return put_str("cl_khr_fp64", cap, v, ret);
        case CL_DEVICE_MAX_WORK_GROUP_SIZE: { size_t s = MAX_WG; return put(&s, sizeof(s), cap, v, ret); }
        case CL_DEVICE_MAX_WORK_ITEM_DIMENSIONS: { cl_uint u = 3; return put(&u, sizeof(u), cap, v, ret); }
        case CL_DEVICE_MAX_WORK_ITEM_SIZES: { size_t s[3] = {MAX_WG, MAX_WG, MAX_WG}; return put(s, sizeof(s), cap, v, ret); }
        case CL_DEVICE_MAX_COMPUTE_UNITS: { cl_uint u = 1; return put(&u, sizeof(u), cap, v, ret); }
        default: { cl_ulong z = 1ull << 30; return put(&z, cap && cap < sizeof(z) ? cap : sizeof(z), cap, v, ret); }
    }
}
cl_int clRetainDevice(cl_device_id) { return CL_SUCCESS; }
cl_int clReleaseDevice(cl_device_id) { return CL_SUCCESS; }

cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *, void(CL_CALLBACK *)(const char *, const void *, size_t, void *),
                           void *, cl_int *err) {
    if (err) *err = CL_SUCCESS;
    return new _cl_context{1};
}
cl_int clRetainContext(cl_context c) { c->refs++; return CL_SUCCESS; }
cl_int clReleaseContext(cl_context c) { if (--c->refs == 0) delete c; return CL_SUCCESS; }
cl_int clGetContextInfo(cl_context, cl_context_info name, size_t cap, void *v, size_t *ret) {
    if (name == CL_CONTEXT_DEVICES) { cl_device_id d = &g_device; return put(&d, sizeof(d), cap, v, ret); }
    cl_uint one = 1;
    return put(&one, sizeof(one), cap, v, ret);
}
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *err) {
    if (err) *err = CL_SUCCESS;
    return new _cl_command_queue{1};
}
cl_int clRetainCommandQueue(cl_command_queue q) { q->refs++; return CL_SUCCESS; }
cl_int clReleaseCommandQueue(cl_command_queue q) { if (--q->refs == 0) delete q; return CL_SUCCESS; }
cl_int clFinish(cl_command_queue) { return CL_SUCCESS; }
cl_int clFlush(cl_command_queue) { return CL_SUCCESS; }

cl_mem clCreateBuffer(cl_context, cl_mem_flags flags, size_t size, void *host, cl_int *err) {
    _cl_mem *m = new _cl_mem{nullptr, size, 1};
    m->host = calloc(size ? size : 1, 1);
    if ((flags & (CL_MEM_COPY_HOST_PTR | CL_MEM_USE_HOST_PTR)) && host) memcpy(m->host, host, size);
    if (err) *err = CL_SUCCESS;
    return m;
}
cl_int clRetainMemObject(cl_mem m) { m->refs++; return CL_SUCCESS; }
cl_int clReleaseMemObject(cl_mem m) { if (--m->refs == 0) { free(m->host); delete m; } return CL_SUCCESS; }

cl_program clCreateProgramWithSource(cl_context, cl_uint count, const char **strings, const size_t *lengths, cl_int *err) {
    _cl_program *p = new _cl_program();
    p->refs = 1; p->dl = nullptr; p->built = false; p->has_barrier = false;
    for (cl_uint i = 0; i < count; i++) {
        std::string s = (lengths && lengths[i]) ? std::string(strings[i], lengths[i]) : std::string(strings[i]);
        while (!s.empty() && s.back() == '\0') s.pop_back();  // the reference passes length()+1
        p->src += s;
        p->src += "\n";
    }
    if (err) *err = CL_SUCCESS;
    return p;
}
cl_int clRetainProgram(cl_program p) { p->refs++; return CL_SUCCESS; }
cl_int clReleaseProgram(cl_program p) { if (--p->refs == 0) delete p; return CL_SUCCESS; }

static uint64_t fnv(const std::string &s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}
static void replace_all(std::string &s, const std::string &a, const std::string &b) {
    for (size_t pos = 0; (pos = s.find(a, pos)) != std::string::npos; pos += b.size()) s.replace(pos, a.size(), b);
}

cl_int clBuildProgram(cl_program p, cl_uint, const cl_device_id *, const char *, void(CL_CALLBACK *)(cl_program, void *), void *) {
    std::string body = p->src;
    if (body.size() >= 3 && (unsigned char)body[0] == 0xEF) body.erase(0, 3);  // UTF-8 BOM
    for (const char *t : {"float2", "float4", "double2"}) replace_all(body, std::string("(") + t + ")(", std::string("mk_") + t + "(");
    p->has_barrier = body.find("barrier(") != std::string::npos;
    // kernel signatures -> argument-unpacking entries
    std::regex sig(R"(__kernel\s+void\s+(\w+)\s*\(([^)]*)\))");
    std::string wrappers;
    for (auto it = std::sregex_iterator(body.begin(), body.end(), sig); it != std::sregex_iterator(); ++it) {
        KInfo k;
        k.name = (*it)[1];
        std::stringstream ps((*it)[2].str());
        std::string prm, call;
        int idx = 0;
        while (std::getline(ps, prm, ',')) {
            const bool ptr = prm.find('*') != std::string::npos;
            std::string ty = "int";
            for (const char *t : {"double", "float", "size_t", "long", "uint", "unsigned", "int"})
                if (!ptr && std::regex_search(prm, std::regex(std::string("\\b") + t + "\\b"))) { ty = (std::string(t) == "uint") ? "unsigned" : t; break; }
            const int kind = ptr ? 0 : ((ty == "double" || ty == "size_t" || ty == "long") ? 8 : 4);
            k.kinds.push_back(kind);
            if (idx) call += ", ";
            call += ptr ? "*(void**)a[" + std::to_string(idx) + "]" : "*(" + ty + "*)a[" + std::to_string(idx) + "]";
            idx++;
        }
        wrappers += "void " + k.name + "__entry(void **a) { " + k.name + "(" + call + "); }\n";
        p->kernels.push_back(k);
    }
    const char *mode = getenv("MINICL_FP_CONTRACT");
    const std::string contract = (mode && std::string(mode) == "fast") ? "fast" : "off";
    // __local arrays are thread-local statics of the kernel module.  The reference's maxPivotKernel writes one element past
    // the end of its 256-entry __local array when the search window has odd length (localData2[loopLimit], loopLimit = 256):
    // harmless in a GPU's local memory, heap corruption in a TLS block that ends exactly there -- so the module's TLS
    // block is fenced with slack on both sides.
    const std::string guard_head = "static __thread char minicl_tls_head[4096] __attribute__((used));\n";
    const std::string guard_tail = "static __thread char minicl_tls_tail[4096] __attribute__((used));\n";
    const std::string full = std::string("#include \"") + MINICL_PRELUDE + "\"\n" + guard_head + body + "\n" + guard_tail + wrappers;
    const char *cd = getenv("MINICL_CACHE");
    const std::string cache = cd ? cd : "/tmp/minicl_cache";
    mkdir(cache.c_str(), 0755);
    char name[64];
    snprintf(name, sizeof(name), "%016llx", (unsigned long long)fnv(full + contract + "v4"));
    const std::string so = cache + "/k" + name + ".so", cfile = cache + "/k" + name + ".c";
    if (access(so.c_str(), R_OK) != 0) {
        { std::ofstream f(cfile); f << full; }
        const std::string tmp = so + "." + std::to_string(getpid());
        const std::string cmd = "gcc -x c -std=gnu11 -O2 -mfma -fno-toplevel-reorder -ffp-contract=" + contract + " -w -shared -fPIC -o " + tmp + " " + cfile +
                                " -lm 2> " + cfile + ".log";
        if (system(cmd.c_str()) != 0) {
            std::ifstream lf(cfile + ".log");
            std::stringstream ss; ss << lf.rdbuf();
            p->log = ss.str();
            fprintf(stderr, "minicl: kernel build failed:\n%s\n", p->log.c_str());
            return CL_BUILD_PROGRAM_FAILURE;
        }
        rename(tmp.c_str(), so.c_str());
    }
    p->dl = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!p->dl) { p->log = dlerror(); return CL_BUILD_PROGRAM_FAILURE; }
    p->setcur = (setcur_fn)dlsym(p->dl, "minicl_set_cur");
    p->setbar = (setbar_fn)dlsym(p->dl, "minicl_set_barrier");
    p->built = true;
    return CL_SUCCESS;
}
cl_int clGetProgramBuildInfo(cl_program p, cl_device_id, cl_program_build_info name, size_t cap, void *v, size_t *ret) {
    if (name == CL_PROGRAM_BUILD_LOG) return put_str(p->log.c_str(), cap, v, ret);
    cl_build_status st = p->built ? CL_BUILD_SUCCESS : CL_BUILD_ERROR;
    return put(&st, sizeof(st), cap, v, ret);
}
cl_int clGetProgramInfo(cl_program, cl_program_info name, size_t cap, void *v, size_t *ret) {
    if (name == CL_PROGRAM_BINARY_SIZES) { size_t z = 0; return put(&z, sizeof(z), cap, v, ret); }
    if (name == CL_PROGRAM_BINARIES) { if (ret) *ret = sizeof(void *); return CL_SUCCESS; }
    cl_uint one = 1;
    return put(&one, sizeof(one), cap, v, ret);
}

cl_kernel clCreateKernel(cl_program p, const char *name, cl_int *err) {
    for (auto &k : p->kernels)
        if (k.name == name) {
            _cl_kernel *kk = new _cl_kernel();
            kk->refs = 1; kk->prog = p; kk->info = k;
            kk->fn = (entry_fn)dlsym(p->dl, (k.name + "__entry").c_str());
            kk->argbytes.resize(k.kinds.size());
            kk->argptr.resize(k.kinds.size(), nullptr);
            if (err) *err = kk->fn ? CL_SUCCESS : CL_INVALID_KERNEL_NAME;
            return kk;
        }
    if (err) *err = CL_INVALID_KERNEL_NAME;
    return nullptr;
}
cl_int clRetainKernel(cl_kernel k) { k->refs++; return CL_SUCCESS; }
cl_int clReleaseKernel(cl_kernel k) { if (--k->refs == 0) delete k; return CL_SUCCESS; }
cl_int clSetKernelArg(cl_kernel k, cl_uint i, size_t size, const void *value) {
    if (i >= k->info.kinds.size()) return CL_INVALID_ARG_INDEX;
    auto &b = k->argbytes[i];
    if (k->info.kinds[i] == 0) {  // buffer: value points at a cl_mem handle
        cl_mem m = value ? *(const cl_mem *)value : nullptr;
        void *hp = m ? m->host : nullptr;
        b.assign((unsigned char *)&hp, (unsigned char *)&hp + sizeof(hp));
    } else {
        b.assign(8, 0);
        memcpy(b.data(), value, size < 8 ? size : 8);
    }
    return CL_SUCCESS;
}
cl_int clGetKernelWorkGroupInfo(cl_kernel, cl_device_id, cl_kernel_work_group_info name, size_t cap, void *v, size_t *ret) {
    if (name == CL_KERNEL_WORK_GROUP_SIZE || name == CL_KERNEL_PREFERRED_WORK_GROUP_SIZE_MULTIPLE) {
        size_t s = (name == CL_KERNEL_WORK_GROUP_SIZE) ? 256 : 32;
        return put(&s, sizeof(s), cap, v, ret);
    }
    cl_ulong z = 0;
    return put(&z, sizeof(z), cap, v, ret);
}

// ---- fibers for work-groups that use barrier() ------------------------------------------------
struct Fiber { ucontext_t ctx; bool done; WI wi; };
static __thread ucontext_t t_sched;
static __thread Fiber *t_cur;
static __thread entry_fn t_fn;
static __thread void **t_args;
static void fiber_barrier() { swapcontext(&t_cur->ctx, &t_sched); }
static void fiber_main() {
    t_fn(t_args);
    t_cur->done = true;
    swapcontext(&t_cur->ctx, &t_sched);
}

cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel k, cl_uint dim, const size_t *off, const size_t *gsz, const size_t *lsz,
                              cl_uint, const cl_event *, cl_event *ev) {
    if (dim < 1 || dim > 3) return CL_INVALID_WORK_DIMENSION;
    const bool ragged_ok = getenv("MINICL_RAGGED") && getenv("MINICL_RAGGED")[0] == '1';
    size_t G[3] = {1, 1, 1}, L[3] = {1, 1, 1}, O[3] = {0, 0, 0}, NG[3] = {1, 1, 1};
    for (cl_uint d = 0; d < dim; d++) { G[d] = gsz[d]; O[d] = off ? off[d] : 0; }
    if (lsz) {
        size_t tot = 1;
        for (cl_uint d = 0; d < dim; d++) {
            L[d] = lsz[d]; tot *= L[d];
            if (L[d] == 0) return CL_INVALID_WORK_GROUP_SIZE;
            if (G[d] % L[d] != 0 && !ragged_ok) return CL_INVALID_WORK_GROUP_SIZE;  // OpenCL 1.2: must divide
        }
        if (tot > MAX_WG) return CL_INVALID_WORK_GROUP_SIZE;
    } else {  // implementation-chosen: one group if it fits, else the largest divisor <= 256 in dimension 0
        size_t l0 = G[0] <= 256 ? G[0] : 256;
        while (l0 > 1 && G[0] % l0 != 0) l0--;
        L[0] = l0 ? l0 : 1;
    }
    for (cl_uint d = 0; d < dim; d++) NG[d] = (G[d] + L[d] - 1) / L[d];
    for (auto &b : k->argbytes) if (b.empty()) return CL_INVALID_KERNEL_ARGS;
    std::vector<void *> args(k->argbytes.size());
    for (size_t i = 0; i < args.size(); i++) args[i] = k->argbytes[i].data();
    const auto t0 = std::chrono::steady_clock::now();
    _cl_program *p = k->prog;
    WI base;
    memset(&base, 0, sizeof(base));
    base.dim = dim;
    for (int d = 0; d < 3; d++) { base.gsz[d] = G[d]; base.lsz[d] = L[d]; base.off[d] = O[d]; base.ngrp[d] = NG[d]; }

    if (!p->has_barrier) {
#pragma omp parallel for schedule(static) collapse(2)
        for (long long z = 0; z < (long long)G[2]; z++)
            for (long long y = 0; y < (long long)G[1]; y++) {
                WI w = base;
                p->setcur(&w);
                for (size_t x = 0; x < G[0]; x++) {
                    const size_t id[3] = {x, (size_t)y, (size_t)z};
                    for (int d = 0; d < 3; d++) { w.gid[d] = id[d] + O[d]; w.lid[d] = id[d] % L[d]; w.grp[d] = id[d] / L[d]; }
                    k->fn(args.data());
                }
            }
    } else {
        p->setbar(fiber_barrier);
        const size_t lsize = L[0] * L[1] * L[2];
        const size_t STK = 64 * 1024;
        std::vector<Fiber> fibers(lsize);
        std::vector<char> stacks(lsize * STK);
        for (size_t gz = 0; gz < NG[2]; gz++)
            for (size_t gy = 0; gy < NG[1]; gy++)
                for (size_t gx = 0; gx < NG[0]; gx++) {
                    const size_t g[3] = {gx, gy, gz};
                    size_t nf = 0;
                    for (size_t lz = 0; lz < L[2]; lz++)
                        for (size_t ly = 0; ly < L[1]; ly++)
                            for (size_t lx = 0; lx < L[0]; lx++) {
                                const size_t l[3] = {lx, ly, lz};
                                bool inside = true;
                                WI w = base;
                                for (int d = 0; d < 3; d++) {
                                    const size_t id = g[d] * L[d] + l[d];
                                    if (id >= G[d]) inside = false;  // ragged last group (lenient mode only)
                                    w.gid[d] = id + O[d]; w.lid[d] = l[d]; w.grp[d] = g[d];
                                }
                                if (!inside) continue;
                                Fiber &f = fibers[nf];
                                f.wi = w; f.done = false;
                                getcontext(&f.ctx);
                                f.ctx.uc_stack.ss_sp = stacks.data() + nf * STK;
                                f.ctx.uc_stack.ss_size = STK;
                                f.ctx.uc_link = nullptr;
                                makecontext(&f.ctx, fiber_main, 0);
                                nf++;
                            }
                    t_fn = k->fn; t_args = args.data();
                    size_t alive = nf;
                    while (alive) {
                        alive = 0;
                        for (size_t i = 0; i < nf; i++) {
                            if (fibers[i].done) continue;
                            t_cur = &fibers[i];
                            p->setcur(&fibers[i].wi);
                            swapcontext(&t_sched, &fibers[i].ctx);
                            if (!fibers[i].done) alive++;
                        }
                    }
                }
    }
    if (ev) {
        const auto t1 = std::chrono::steady_clock::now();
        _cl_event *e = new _cl_event{1, 0, 0};
        e->t0 = (cl_ulong)std::chrono::duration_cast<std::chrono::nanoseconds>(t0.time_since_epoch()).count();
        e->t1 = (cl_ulong)std::chrono::duration_cast<std::chrono::nanoseconds>(t1.time_since_epoch()).count();
        *ev = e;
    }
    return CL_SUCCESS;
}

cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem m, cl_bool, size_t off, size_t size, void *dst, cl_uint, const cl_event *, cl_event *ev) {
    if (off + size > m->size) return CL_INVALID_VALUE;
    memcpy(dst, (char *)m->host + off, size);
    if (ev) *ev = new _cl_event{1, 0, 0};
    return CL_SUCCESS;
}
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem m, cl_bool, size_t off, size_t size, const void *src, cl_uint, const cl_event *, cl_event *ev) {
    if (off + size > m->size) return CL_INVALID_VALUE;
    memcpy((char *)m->host + off, src, size);
    if (ev) *ev = new _cl_event{1, 0, 0};
    return CL_SUCCESS;
}
cl_int clWaitForEvents(cl_uint, const cl_event *) { return CL_SUCCESS; }
cl_int clRetainEvent(cl_event e) { e->refs++; return CL_SUCCESS; }
cl_int clReleaseEvent(cl_event e) { if (--e->refs == 0) delete e; return CL_SUCCESS; }
cl_int clGetEventProfilingInfo(cl_event e, cl_profiling_info name, size_t cap, void *v, size_t *ret) {
    cl_ulong t = (name == CL_PROFILING_COMMAND_END) ? e->t1 : e->t0;
    return put(&t, sizeof(t), cap, v, ret);
}
cl_int clGetEventInfo(cl_event, cl_event_info, size_t cap, void *v, size_t *ret) {
    cl_int st = CL_COMPLETE;
    return put(&st, sizeof(st), cap, v, ret);
}

}  // extern "C"
