/*
 * ref_driver.c -- TEST/BENCH INFRASTRUCTURE ONLY.
 *
 * Times the REFERENCE library (oracle/_ref/libblingfiretokdll.so, built from the reference's own
 * sources) on a CSR batch with N host threads: one shared model handle, the threads pinned one per
 * CPU and pulling blocks of 256 documents from a shared counter, no Python in the loop (SURVEY 8d "CPU baseline beside it").  Used by bench.py's
 * cpu_baseline leg and by `bench.py --impl reference`; never by the product.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef void* (*load_fn)(const char*);
typedef int (*free_fn)(void*);
typedef int (*ids_fn)(void*, const char*, int, int32_t*, int, int);

typedef struct {
    ids_fn f; void* model; const char* text; const int64_t* offs; int64_t ndocs;
    int max_ids, unk, tid, nthreads; int64_t tokens; int32_t* counts;
    uint64_t* digests;       /* optional: FNV-1a-64 of every document's ids (SURVEY 8c recipe) */
    atomic_llong* next;      /* shared block counter: the threads pull blocks of documents (ragged lengths balance) */
    int cpu;                 /* CPU to run on, -1 = wherever the scheduler puts the thread */
} job_t;

enum { kBlockDocs = 256 };

static void* worker(void* a) {
    job_t* j = (job_t*)a;
    if (j->cpu >= 0) {
        cpu_set_t set; CPU_ZERO(&set); CPU_SET(j->cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)(j->max_ids > 0 ? j->max_ids : 1));
    int64_t tot = 0;
    for (;;) {
        const int64_t b = (int64_t)atomic_fetch_add(j->next, (long long)kBlockDocs);
        if (b >= j->ndocs) break;
        const int64_t e = b + kBlockDocs < j->ndocs ? b + kBlockDocs : j->ndocs;
        for (int64_t d = b; d < e; ++d) {
            const int c = j->f(j->model, j->text + j->offs[d], (int)(j->offs[d + 1] - j->offs[d]), ids, j->max_ids, j->unk);
            if (j->counts) j->counts[d] = c;
            if (j->digests) {
                uint64_t h = 0xcbf29ce484222325ull;
                const int n = c < j->max_ids ? c : j->max_ids;
                for (int k = 0; k < n; ++k) { h ^= (uint32_t)ids[k]; h *= 0x100000001b3ull; }
                j->digests[d] = h;
            }
            tot += c;
        }
    }
    j->tokens = tot;
    free(ids);
    return NULL;
}

static void* g_lib = NULL;
static void* g_model = NULL;
static char g_model_path[4096];

/* Returns elapsed seconds (wall clock of the threaded region), or -1 on error.
 * counts (optional) receives the per-document return values. */
double ref_digest_batch(const char* lib_path, const char* model_path, const char* text, const int64_t* offs,
                        int64_t ndocs, int max_ids, int unk, int threads, int64_t* tokens, int32_t* counts, uint64_t* digests);

double ref_time_batch(const char* lib_path, const char* model_path, const char* text, const int64_t* offs,
                      int64_t ndocs, int max_ids, int unk, int threads, int64_t* tokens, int32_t* counts) {
    return ref_digest_batch(lib_path, model_path, text, offs, ndocs, max_ids, unk, threads, tokens, counts, NULL);
}

/* same, and the per-document digests of the reference's ids (full-size parity without an id matrix) */
double ref_digest_batch(const char* lib_path, const char* model_path, const char* text, const int64_t* offs,
                        int64_t ndocs, int max_ids, int unk, int threads, int64_t* tokens, int32_t* counts, uint64_t* digests) {
    if (!g_lib) { g_lib = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL); if (!g_lib) return -1.0; }
    load_fn load = (load_fn)dlsym(g_lib, "LoadModel");
    ids_fn f = (ids_fn)dlsym(g_lib, "TextToIds");
    if (!load || !f) return -1.0;
    if (!g_model || strcmp(g_model_path, model_path) != 0) {
        g_model = load(model_path);
        if (!g_model) return -1.0;
        strncpy(g_model_path, model_path, sizeof(g_model_path) - 1);
    }
    if (threads < 1) threads = 1;
    if (threads > 512) threads = 512;
    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    /* one thread per allowed CPU, in order, while there are enough CPUs; otherwise unpinned */
    cpu_set_t allowed; CPU_ZERO(&allowed);
    int cpus[CPU_SETSIZE], ncpus = 0;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &allowed)) cpus[ncpus++] = c;
    atomic_llong next; atomic_init(&next, 0);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; ++t) {
        /* spread a partial set of threads over the whole machine (both sockets) rather than packing them */
        const int cpu = threads <= ncpus ? cpus[(int)((int64_t)t * ncpus / threads)] : -1;
        job_t j = { f, g_model, text, offs, ndocs, max_ids, unk, t, threads, 0, counts, digests, &next, cpu };
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    int64_t tot = 0;
    for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); tot += jobs[t].tokens; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (tokens) *tokens = tot;
    free(jobs); free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
