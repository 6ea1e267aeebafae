// hrp_comm.cu -- the one exchange of the data-parallel PPO path, fused with its consumer.
//
// Every rank (one process per GPU, same box) holds the gradient of its shard of the minibatch in a buffer that its
// peers can read through NVLink (cudaIpc).  `hrp_clip_adam_step_p2p` is ONE cooperative kernel per optimizer step:
//
//   1. cross-GPU barrier: "my gradient is complete" flags written into every peer, wait for every peer's flag;
//   2. every CTA sums its slice of the W gradients in rank order (peer loads over NVLink; the same order on every
//      rank, so all ranks hold bit-identical sums), keeps the sum locally and accumulates its square;
//   3. grid.sync(), fixed-order total -> clip coefficient (clip_grad_norm_, agent.py:249);
//   4. Adam on the slice (agent.py:252), step counter;
//   5. cross-GPU barrier: "I have finished reading" -- after it a rank may overwrite its gradient buffer.
//
// It replaces all_reduce(grad) + gradnorm + clip + Adam (NCCL launch + 1 kernel) of the NCCL path.  The flags carry
// a monotonically increasing epoch kept in device memory, so the launch can be captured in a CUDA graph and replayed.
// Kernels of different ranks spin on each other: every rank owns its GPU (never run two ranks on one device).
#include <cooperative_groups.h>
#include <math.h>
#include <string.h>

#include "hrp_internal.cuh"

#define HRP_MAX_RANKS 8

struct hrp_comm {
    int world, rank, device;
    long long n;             // floats per gradient
    float *local;            // [grad n | sum n | flags], the IPC-exported allocation
    void *peer_base[HRP_MAX_RANKS];
    float *peer_grad[HRP_MAX_RANKS];
    unsigned *peer_flags[HRP_MAX_RANKS];   // each rank's flag block: ready[HRP_MAX_RANKS], done[HRP_MAX_RANKS]
    unsigned *epoch;         // device: [0] counter of completed steps, [1] timeout bits of the cross-GPU waits
    bool connected;
};

namespace {

constexpr int P2P_MAX_CTAS = 120, P2P_THREADS = 512;

struct PeerTable {
    const float *grad[HRP_MAX_RANKS];
    unsigned *flags[HRP_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// flag block of a rank: ready[q] / done[q] are written by rank q.  The wait is bounded (about fifteen seconds of SM
// clock): a rank that never arrives -- a crashed peer -- must not leave this GPU spinning; the timeout is recorded in
// *status (bit `which`) and the step goes on with whatever the peer buffers hold, for the host to detect.
__device__ __forceinline__ void cross_gpu_barrier(const PeerTable &T, int world, int rank, int which, unsigned epoch,
                                                  unsigned *status)
{
    // called by the first `world` threads of CTA 0
    const int q = threadIdx.x;
    if (q < world) {
        __threadfence_system();
        st_release_sys(T.flags[q] + which * HRP_MAX_RANKS + rank, epoch);
        const unsigned *mine = T.flags[rank] + which * HRP_MAX_RANKS + q;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            __nanosleep(64);
            if (clock64() - t0 > 30000000000ll) { atomicOr(status, 1u << which); break; }
        }
    }
}

__global__ void __launch_bounds__(P2P_THREADS)
clip_adam_p2p_kernel(const PeerTable T, int world, int rank, unsigned *__restrict__ epoch_dev, float *__restrict__ gsum,
                     float *__restrict__ p, float *__restrict__ m, float *__restrict__ v, int32_t *__restrict__ step,
                     long long n, double lr, double beta1, double beta2, double eps, float max_norm,
                     float *__restrict__ part)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ float red[P2P_THREADS / 32];
    __shared__ float s_coef, s_step_size, s_bc2_sqrt;
    const unsigned epoch = *epoch_dev + 1u;
    const int k = step[0] + 1;
    if (blockIdx.x == 0) cross_gpu_barrier(T, world, rank, 0, epoch, epoch_dev + 1);   // every gradient is complete
    grid.sync();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f;
    for (long long i = i0; i < n; i += stride) {
        float g = 0.f;
        for (int q = 0; q < world; ++q) g += __ldcg(T.grad[q] + i);      // rank order: identical on every rank
        gsum[i] = g;
        s = fmaf(g, g, s);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(HRP_FULL, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < P2P_THREADS / 32; ++i) t += red[i];
        part[blockIdx.x] = t;
    }
    grid.sync();
    if (blockIdx.x == 0) cross_gpu_barrier(T, world, rank, 1, epoch, epoch_dev + 1);   // nobody reads my gradient any more
    if (threadIdx.x < 32) {
        float t = 0.f;
        for (int i = threadIdx.x * 4; i < min((int)gridDim.x, threadIdx.x * 4 + 4); ++i) t += __ldcg(part + i);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(HRP_FULL, t, d);
        if (threadIdx.x == 0) {
            float total_norm = sqrtf(t);
            float c = max_norm / (total_norm + 1e-6f);
            s_coef = max_norm > 0.f ? fminf(c, 1.f) : 1.f;
            double bc1 = 1.0 - pow(beta1, (double)k), bc2 = 1.0 - pow(beta2, (double)k);
            s_step_size = (float)(lr / bc1);
            s_bc2_sqrt = (float)sqrt(bc2);
        }
    }
    __syncthreads();
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), ep = (float)eps;
    const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    for (long long i = i0; i < n; i += stride) {
        float gi = gsum[i] * coef;
        float mi = m[i] + w1 * (gi - m[i]);
        float vi = v[i] * b2 + w2 * gi * gi;
        m[i] = mi; v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + ep;
        p[i] = p[i] - step_size * (mi / denom);
    }
    grid.sync();   // every CTA has read step[0] / *epoch_dev and passed both barriers
    if (blockIdx.x == 0 && threadIdx.x == 0) { step[0] = k; *epoch_dev = epoch; }
}

inline size_t comm_floats(long long n) { return (size_t)(2 * ((n + 31) / 32 * 32) + 64); }

}  // namespace

extern "C" {

int hrp_comm_create(int32_t world, int32_t rank, int64_t n, int32_t device, hrp_comm **out, void *handle_out64)
{
    if (!out || !handle_out64 || world < 1 || world > HRP_MAX_RANKS || rank < 0 || rank >= world || n < 1) {
        hrp_set_error("hrp_comm_create: bad arguments (world <= %d)", HRP_MAX_RANKS);
        return -1;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    HRP_CUDA_OK(cudaSetDevice(device));
    hrp_comm *c = new hrp_comm();
    c->world = world; c->rank = rank; c->device = device; c->n = n;
    size_t bytes = comm_floats(n) * sizeof(float);
    cudaError_t ce = cudaMalloc(&c->local, bytes);
    if (ce == cudaSuccess) ce = cudaMemset(c->local, 0, bytes);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->epoch, 2 * sizeof(unsigned));
    if (ce == cudaSuccess) ce = cudaMemset(c->epoch, 0, 2 * sizeof(unsigned));
    cudaIpcMemHandle_t hd;
    if (ce == cudaSuccess) ce = cudaIpcGetMemHandle(&hd, c->local);
    if (ce != cudaSuccess) {
        hrp_set_error("hrp_comm_create: %s", cudaGetErrorString(ce));
        cudaFree(c->local); cudaFree(c->epoch); delete c;
        return -2;
    }
    memcpy(handle_out64, &hd, 64);
    HRP_CUDA_OK(cudaDeviceSynchronize());
    *out = c;
    return 0;
}

/* handles: world x 64 bytes, rank-major, as gathered from every rank's hrp_comm_create */
int hrp_comm_connect(hrp_comm *c, const void *handles)
{
    if (!c || !handles) { hrp_set_error("hrp_comm_connect: null argument"); return -1; }
    HRP_CUDA_OK(cudaSetDevice(c->device));
    const size_t pad = (size_t)((c->n + 31) / 32 * 32);
    for (int q = 0; q < c->world; ++q) {
        void *base = c->local;
        if (q != c->rank) {
            cudaIpcMemHandle_t hd;
            memcpy(&hd, (const char *)handles + 64 * q, 64);
            HRP_CUDA_OK(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
        }
        c->peer_base[q] = base;
        c->peer_grad[q] = (float *)base;
        c->peer_flags[q] = (unsigned *)((float *)base + 2 * pad);
    }
    c->connected = true;
    return 0;
}

/* device pointer of this rank's gradient buffer (n floats): hrp_ppo_loss_grad writes it */
float *hrp_comm_grad(hrp_comm *c) { return c ? c->local : nullptr; }

int hrp_clip_adam_step_p2p(hrp_comm *c, float *params, float *exp_avg, float *exp_avg_sq, int32_t *step, double lr,
                           double beta1, double beta2, double eps, float max_grad_norm, float *scratch, void *stream)
{
    if (!c || !c->connected || !params || !exp_avg || !exp_avg_sq || !step || !scratch) {
        hrp_set_error("hrp_clip_adam_step_p2p: bad arguments (hrp_comm_connect first)");
        return -1;
    }
    static int max_ctas_dev[HRP_MAX_DEVICES] = {0};
    if (c->device < 0 || c->device >= HRP_MAX_DEVICES) { hrp_set_error("device index %d not supported", c->device); return -1; }
    int &max_ctas = max_ctas_dev[c->device];
    if (max_ctas == 0) {
        int sms = 0, per_sm = 0;
        HRP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        HRP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, clip_adam_p2p_kernel, P2P_THREADS, 0));
        max_ctas = sms * per_sm < P2P_MAX_CTAS ? sms * per_sm : P2P_MAX_CTAS;
        if (max_ctas < 1) { hrp_set_error("hrp_clip_adam_step_p2p: kernel does not fit the device"); return -2; }
    }
    PeerTable T;
    for (int q = 0; q < HRP_MAX_RANKS; ++q) {
        T.grad[q] = q < c->world ? c->peer_grad[q] : nullptr;
        T.flags[q] = q < c->world ? c->peer_flags[q] : nullptr;
    }
    long long n = c->n;
    int ctas = (int)((n + P2P_THREADS - 1) / P2P_THREADS);
    if (ctas > max_ctas) ctas = max_ctas;
    int world = c->world, rank = c->rank;
    unsigned *epoch = c->epoch;
    float *gsum = c->local + (size_t)((n + 31) / 32 * 32);
    void *args[] = {&T, &world, &rank, &epoch, &gsum, &params, &exp_avg, &exp_avg_sq, &step, &n, &lr, &beta1, &beta2, &eps,
                    &max_grad_norm, &scratch};
    HRP_CUDA_OK(cudaLaunchCooperativeKernel((const void *)clip_adam_p2p_kernel, dim3(ctas), dim3(P2P_THREADS), args, 0,
                                            (cudaStream_t)stream));
    return 0;
}

/* synchronises the device; 0 = every cross-GPU wait so far was answered, -4 = a peer did not arrive within the
 * timeout of some step (its results are not to be trusted) */
int hrp_comm_status(hrp_comm *c)
{
    if (!c) { hrp_set_error("hrp_comm_status: null argument"); return -1; }
    HRP_CUDA_OK(cudaSetDevice(c->device));
    unsigned st[2] = {0, 0};
    HRP_CUDA_OK(cudaMemcpy(st, c->epoch, sizeof(st), cudaMemcpyDeviceToHost));
    if (st[1] != 0) { hrp_set_error("hrp_comm: a peer rank did not reach the gradient exchange (wait bits %u) after %u steps", st[1], st[0]); return -4; }
    return 0;
}

int hrp_comm_destroy(hrp_comm *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < c->world; ++q)
        if (c->connected && q != c->rank && c->peer_base[q]) cudaIpcCloseMemHandle(c->peer_base[q]);
    cudaFree(c->local);
    cudaFree(c->epoch);
    delete c;
    return 0;
}

}  // extern "C"
