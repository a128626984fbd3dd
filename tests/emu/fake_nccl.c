/* TEST INFRASTRUCTURE ONLY: the six NCCL entry points csrc/api.cu binds (dlopen), implemented over POSIX shared memory for the
 * CPU-emulated build, so that the multi-rank paths (row-sharded control-variate pass, gathered modes, sharded mode finder, sharded
 * dataset) run as real multi-process jobs in the GPU-less container.  "Device" pointers of the emulated build are host pointers, the
 * emulated streams are synchronous, so every collective completes inline.  Only sum reductions of float32 / float64. */
#define _GNU_SOURCE
#include <fcntl.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#define ID_BYTES 128
#define MAX_RANKS 16
#define SLOT_BYTES (16u << 20)

typedef struct { char internal[ID_BYTES]; } ncclUniqueId;
typedef struct {
    atomic_int arrived;
    atomic_int generation;
    atomic_int attached;
} Header;
typedef struct FakeComm {
    int rank, nranks;
    char name[ID_BYTES];
    Header* hdr;
    unsigned char* slots;   /* [nranks][SLOT_BYTES] */
    size_t bytes;
} FakeComm;
typedef FakeComm* ncclComm_t;

static const char* g_last = "no error";

static void barrier(FakeComm* c) {
    const int gen = atomic_load(&c->hdr->generation);
    if (atomic_fetch_add(&c->hdr->arrived, 1) + 1 == c->nranks) {
        atomic_store(&c->hdr->arrived, 0);
        atomic_fetch_add(&c->hdr->generation, 1);
    } else {
        long spins = 0;
        while (atomic_load(&c->hdr->generation) == gen) {
            if (++spins > 60000000L) { fprintf(stderr, "fake_nccl: rank %d waited too long in a collective\n", c->rank); abort(); }
            sched_yield();
        }
    }
}

static size_t elem_size(int dtype) { return dtype == 7 ? 4 : (dtype == 8 ? 8 : 0); }   /* ncclFloat32 = 7, ncclFloat64 = 8 */

int ncclGetUniqueId(ncclUniqueId* id) {
    memset(id, 0, sizeof *id);
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    snprintf(id->internal, ID_BYTES, "/sgm_fake_nccl_%d_%ld_%ld", (int)getpid(), (long)ts.tv_sec, (long)ts.tv_nsec);
    return 0;
}

int ncclCommInitRank(ncclComm_t* out, int nranks, ncclUniqueId id, int rank) {
    if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) { g_last = "bad rank / world size"; return 4; }
    FakeComm* c = (FakeComm*)calloc(1, sizeof *c);
    c->rank = rank; c->nranks = nranks;
    snprintf(c->name, sizeof c->name, "%s", id.internal);
    c->bytes = 4096 + (size_t)nranks * SLOT_BYTES;
    int fd = shm_open(c->name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) { g_last = "shm_open failed"; free(c); return 2; }
    if (ftruncate(fd, (off_t)c->bytes) != 0) { g_last = "ftruncate failed"; close(fd); free(c); return 2; }
    void* p = mmap(NULL, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { g_last = "mmap failed"; free(c); return 2; }
    c->hdr = (Header*)p;               /* a fresh shm object is zero-filled: counters start at 0 */
    c->slots = (unsigned char*)p + 4096;
    atomic_fetch_add(&c->hdr->attached, 1);
    long spins = 0;
    while (atomic_load(&c->hdr->attached) < nranks) { if (++spins > 2000000000L) { g_last = "peers never attached"; return 2; } sched_yield(); }
    *out = c;
    return 0;
}

int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, ncclComm_t c, void* stream) {
    (void)stream;
    const size_t es = elem_size(dtype);
    if (!es || op != 0) { g_last = "only sum of float32 / float64"; return 4; }
    if (count * es > SLOT_BYTES) { g_last = "message larger than the fake transport's slot"; return 4; }
    memcpy(c->slots + (size_t)c->rank * SLOT_BYTES, send, count * es);
    barrier(c);
    for (size_t i = 0; i < count; ++i) {
        if (es == 8) { double s = 0; for (int r = 0; r < c->nranks; ++r) s += ((const double*)(c->slots + (size_t)r * SLOT_BYTES))[i]; ((double*)recv)[i] = s; }
        else { float s = 0; for (int r = 0; r < c->nranks; ++r) s += ((const float*)(c->slots + (size_t)r * SLOT_BYTES))[i]; ((float*)recv)[i] = s; }
    }
    barrier(c);
    return 0;
}

int ncclAllGather(const void* send, void* recv, size_t sendcount, int dtype, ncclComm_t c, void* stream) {
    (void)stream;
    const size_t es = elem_size(dtype);
    if (!es) { g_last = "only float32 / float64"; return 4; }
    if (sendcount * es > SLOT_BYTES) { g_last = "message larger than the fake transport's slot"; return 4; }
    memcpy(c->slots + (size_t)c->rank * SLOT_BYTES, send, sendcount * es);
    barrier(c);
    for (int r = 0; r < c->nranks; ++r) memcpy((unsigned char*)recv + (size_t)r * sendcount * es, c->slots + (size_t)r * SLOT_BYTES, sendcount * es);
    barrier(c);
    return 0;
}

int ncclReduceScatter(const void* send, void* recv, size_t recvcount, int dtype, int op, ncclComm_t c, void* stream) {
    (void)stream;
    const size_t es = elem_size(dtype);
    if (!es || op != 0) { g_last = "only sum of float32 / float64"; return 4; }
    if (recvcount * c->nranks * es > SLOT_BYTES) { g_last = "message larger than the fake transport's slot"; return 4; }
    memcpy(c->slots + (size_t)c->rank * SLOT_BYTES, send, recvcount * c->nranks * es);
    barrier(c);
    for (size_t i = 0; i < recvcount; ++i) {
        const size_t j = (size_t)c->rank * recvcount + i;
        if (es == 8) { double s = 0; for (int r = 0; r < c->nranks; ++r) s += ((const double*)(c->slots + (size_t)r * SLOT_BYTES))[j]; ((double*)recv)[i] = s; }
        else { float s = 0; for (int r = 0; r < c->nranks; ++r) s += ((const float*)(c->slots + (size_t)r * SLOT_BYTES))[j]; ((float*)recv)[i] = s; }
    }
    barrier(c);
    return 0;
}

int ncclCommDestroy(ncclComm_t c) {
    if (!c) return 0;
    if (c->rank == 0) shm_unlink(c->name);
    munmap((void*)c->hdr, c->bytes);
    free(c);
    return 0;
}

const char* ncclGetErrorString(int rc) { (void)rc; return g_last; }
