/* boundary.c -- a plain C caller of the C ABI, written the way the cgo shim of INTEGRATION.md uses it:
 *   - polynomials live in cudaMallocManaged memory (the Go side keeps []uint64 views of the rows, ring/poly.go:19-43),
 *     are filled by the HOST, transformed by the library and read back by the host after lgpu_sync;
 *   - the caller's own root tables are pushed through lgpu_ring_set_roots (ring/ntt.go:38-44);
 *   - 8 threads hammer two entry points on 8 streams at once (the reference's ring methods are safe for concurrent use,
 *     ring/ring.go:184-186) and must reproduce the single-threaded words.
 * Prints "boundary ok" and exits 0 on success. Built and run by tests/test_gpu_cabi_boundary.py. */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "lattigo_b200.h"

#define LOGN 12
#define N (1 << LOGN)
#define NQ 3
#define NP 2
static const uint64_t Q[NQ] = {0x1fffffffffe00001ULL, 0x1fffffffffc80001ULL, 0x1fffffffffb40001ULL}; /* ring/test_params.go Qi60 */
static const uint64_t P[NP] = {0x1ffffffff6c80001ULL, 0x1ffffffff6140001ULL};                        /* Pi60 */

#define CHECK(x) do { if ((x) != 0) { fprintf(stderr, "FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, lgpu_last_error()); exit(1); } } while (0)
#define CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA FAIL %s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
#define EXPECT(c, msg) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, msg); exit(1); } } while (0)

static uint64_t lcg(uint64_t* s) { *s = *s * 6364136223846793005ULL + 1442695040888963407ULL; return *s; }
static void fill(uint64_t* p, const uint64_t* mods, int rows, uint64_t seed) {
    for (int i = 0; i < rows; i++) for (int j = 0; j < N; j++) p[(size_t)i * N + j] = lcg(&seed) % mods[i];
}

static lgpu_ctx* ctx;

struct job { int id; uint64_t* x; uint64_t* y; uint64_t* z; uint64_t* want_y; int ok; };
static void* worker(void* arg) {
    struct job* j = (struct job*)arg;
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { j->ok = 0; return NULL; }
    j->ok = 1;
    for (int it = 0; it < 25 && j->ok; it++) {
        if (lgpu_ntt(ctx, LGPU_RING_Q, NQ - 1, j->x, j->y, 0, 1, 0, st)) j->ok = 0;
        if (lgpu_vecop(ctx, LGPU_RING_Q, NQ - 1, LGPU_OP_ADD, j->y, j->y, j->z, NULL, NULL, 1, 0, st)) j->ok = 0;   /* z = 2y */
        if (lgpu_intt(ctx, LGPU_RING_Q, NQ - 1, j->z, j->z, 0, 1, 0, st)) j->ok = 0;                                  /* z = 2x */
        if (lgpu_sync(ctx, st)) j->ok = 0;
        if (memcmp(j->y, j->want_y, sizeof(uint64_t) * NQ * N)) j->ok = 0;
        for (int i = 0; i < NQ && j->ok; i++)
            for (int k = 0; k < N; k++) {
                uint64_t t = 2 * j->x[(size_t)i * N + k]; if (t >= Q[i]) t -= Q[i];
                if (j->z[(size_t)i * N + k] != t) { j->ok = 0; break; }
            }
    }
    cudaStreamDestroy(st);
    return NULL;
}

int main(void) {
    CHECK(lgpu_create(&ctx, 0, LOGN, LGPU_RING_STANDARD, Q, NQ, P, NP));
    const size_t words = (size_t)NQ * N;
    uint64_t *x, *y, *z;
    CUDA(cudaMallocManaged((void**)&x, words * 8, cudaMemAttachGlobal));
    CUDA(cudaMallocManaged((void**)&y, words * 8, cudaMemAttachGlobal));
    CUDA(cudaMallocManaged((void**)&z, words * 8, cudaMemAttachGlobal));

    /* 1. managed polynomials: host writes, library transforms, host reads after lgpu_sync */
    fill(x, Q, NQ, 1);
    CHECK(lgpu_ntt(ctx, LGPU_RING_Q, NQ - 1, x, y, 0, 1, 0, NULL));
    CHECK(lgpu_intt(ctx, LGPU_RING_Q, NQ - 1, y, z, 0, 1, 0, NULL));
    CHECK(lgpu_sync(ctx, NULL));
    EXPECT(memcmp(x, z, words * 8) == 0, "INTT(NTT(x)) != x on managed memory");
    int canonical = 1, differs = 0;
    for (int i = 0; i < NQ; i++) for (int k = 0; k < N; k++) { if (y[(size_t)i * N + k] >= Q[i]) canonical = 0; if (y[(size_t)i * N + k] != x[(size_t)i * N + k]) differs = 1; }
    EXPECT(canonical && differs, "NTT output not canonical / not transformed");
    uint64_t* y_ref = (uint64_t*)malloc(words * 8);
    memcpy(y_ref, y, words * 8);

    /* 2. the caller's tables through lgpu_ring_set_roots: same tables -> same words; perturbed table -> different words */
    uint64_t* fwd = (uint64_t*)malloc(sizeof(uint64_t) * N); uint64_t* bwd = (uint64_t*)malloc(sizeof(uint64_t) * N); uint64_t consts[6];
    for (int i = 0; i < NQ; i++) {
        CHECK(lgpu_ring_get_table(ctx, LGPU_RING_Q, i, 0, consts, 6));
        CHECK(lgpu_ring_get_table(ctx, LGPU_RING_Q, i, 1, fwd, N));
        CHECK(lgpu_ring_get_table(ctx, LGPU_RING_Q, i, 2, bwd, N));
        EXPECT(consts[0] == Q[i], "table read-back: modulus");
        CHECK(lgpu_ring_set_roots(ctx, LGPU_RING_Q, i, fwd, bwd, consts[4]));
    }
    CHECK(lgpu_ntt(ctx, LGPU_RING_Q, NQ - 1, x, y, 0, 1, 0, NULL));
    CHECK(lgpu_sync(ctx, NULL));
    EXPECT(memcmp(y, y_ref, words * 8) == 0, "set_roots with the library's own tables changed the transform");
    CHECK(lgpu_ring_get_table(ctx, LGPU_RING_Q, 1, 0, consts, 6));
    CHECK(lgpu_ring_get_table(ctx, LGPU_RING_Q, 1, 1, fwd, N));
    CHECK(lgpu_ring_get_table(ctx, LGPU_RING_Q, 1, 2, bwd, N));
    { uint64_t t = fwd[5]; fwd[5] = fwd[6]; fwd[6] = t; }
    CHECK(lgpu_ring_set_roots(ctx, LGPU_RING_Q, 1, fwd, bwd, consts[4]));
    CHECK(lgpu_ntt(ctx, LGPU_RING_Q, NQ - 1, x, y, 0, 1, 0, NULL));
    CHECK(lgpu_sync(ctx, NULL));
    EXPECT(memcmp(y, y_ref, (size_t)N * 8) == 0 && memcmp(y + N, y_ref + N, (size_t)N * 8) != 0, "caller tables are not the ones the transform uses");
    { uint64_t t = fwd[5]; fwd[5] = fwd[6]; fwd[6] = t; }
    CHECK(lgpu_ring_set_roots(ctx, LGPU_RING_Q, 1, fwd, bwd, consts[4]));
    CHECK(lgpu_ntt(ctx, LGPU_RING_Q, NQ - 1, x, y, 0, 1, 0, NULL));
    CHECK(lgpu_sync(ctx, NULL));
    EXPECT(memcmp(y, y_ref, words * 8) == 0, "restoring the tables did not restore the transform");

    /* 3. key switch on managed memory == key switch on cudaMalloc memory */
    {
        const int rows = NQ + NP, nd = (NQ - 1 + NP) / NP;   /* BaseRNSDecompositionVectorSize(levelQ=NQ-1, levelP=NP-1) */
        const size_t kw = (size_t)nd * 2 * rows * N;
        uint64_t *km, *kd, *c0m, *c1m, *cxd, *c0d, *c1d;
        uint64_t mods[NQ + NP]; memcpy(mods, Q, sizeof(Q)); memcpy(mods + NQ, P, sizeof(P));
        CUDA(cudaMallocManaged((void**)&km, kw * 8, cudaMemAttachGlobal));
        for (int b = 0; b < nd * 2; b++) fill(km + (size_t)b * rows * N, mods, rows, 100 + b);
        CUDA(cudaMallocManaged((void**)&c0m, words * 8, cudaMemAttachGlobal));
        CUDA(cudaMallocManaged((void**)&c1m, words * 8, cudaMemAttachGlobal));
        CUDA(cudaMalloc((void**)&kd, kw * 8)); CUDA(cudaMalloc((void**)&cxd, words * 8)); CUDA(cudaMalloc((void**)&c0d, words * 8)); CUDA(cudaMalloc((void**)&c1d, words * 8));
        CUDA(cudaMemcpy(kd, km, kw * 8, cudaMemcpyDefault)); CUDA(cudaMemcpy(cxd, x, words * 8, cudaMemcpyDefault));
        lgpu_gadget_ct em = {km, NQ - 1, NP - 1, 0, nd, 1, NULL}, ed = {kd, NQ - 1, NP - 1, 0, nd, 1, NULL};
        CHECK(lgpu_gadget_product(ctx, NQ - 1, x, &em, c0m, c1m, 1, 0, 0, NULL));
        CHECK(lgpu_gadget_product(ctx, NQ - 1, cxd, &ed, c0d, c1d, 1, 0, 0, NULL));
        CHECK(lgpu_sync(ctx, NULL));
        uint64_t* h = (uint64_t*)malloc(words * 8);
        CUDA(cudaMemcpy(h, c0d, words * 8, cudaMemcpyDefault)); EXPECT(memcmp(h, c0m, words * 8) == 0, "gadget product: managed vs device memory (c0)");
        CUDA(cudaMemcpy(h, c1d, words * 8, cudaMemcpyDefault)); EXPECT(memcmp(h, c1m, words * 8) == 0, "gadget product: managed vs device memory (c1)");
        free(h);
    }

    /* 4. eight threads, eight streams, two entry points each */
    enum { T = 8 };
    pthread_t th[T]; struct job jobs[T];
    for (int t = 0; t < T; t++) {
        jobs[t].id = t;
        CUDA(cudaMallocManaged((void**)&jobs[t].x, words * 8, cudaMemAttachGlobal));
        CUDA(cudaMallocManaged((void**)&jobs[t].y, words * 8, cudaMemAttachGlobal));
        CUDA(cudaMallocManaged((void**)&jobs[t].z, words * 8, cudaMemAttachGlobal));
        jobs[t].want_y = (uint64_t*)malloc(words * 8);
        fill(jobs[t].x, Q, NQ, 1000 + t);
        CHECK(lgpu_ntt(ctx, LGPU_RING_Q, NQ - 1, jobs[t].x, jobs[t].y, 0, 1, 0, NULL));
        CHECK(lgpu_sync(ctx, NULL));
        memcpy(jobs[t].want_y, jobs[t].y, words * 8);
    }
    for (int t = 0; t < T; t++) pthread_create(&th[t], NULL, worker, &jobs[t]);
    for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < T; t++) EXPECT(jobs[t].ok, "concurrent callers produced different words than the single-threaded run");

    lgpu_destroy(ctx);
    printf("boundary ok\n");
    return 0;
}
