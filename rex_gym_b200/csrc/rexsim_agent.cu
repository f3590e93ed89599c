// rexsim_agent.cu -- device-side agent glue around the step kernel (include/rexsim_agent.h):
//   perform     : StreamingNormalize.transform + ForwardGaussianPolicy (policy MLP, value MLP) + sampling, one fused kernel
//   experience  : StreamingNormalize.update for observations and rewards, one deterministic reduction kernel
//   scans       : discounted_return / lambda_advantage (reference row semantics) and the done-aware time-major GAE
// fp32 throughout (the reference's TF graph is fp32): the MLP is computed on the CUDA cores with a register-tiled
// shared-memory GEMM so results match an fp32 reference to rounding; see DESIGN.md for the roofline of each kernel.
#include "rexsim_kernel.cuh"
#include "../../include/rexsim_agent.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include <type_traits>

namespace rexsim {
void set_error(const char* msg);   // rexsim_capi.cu (message returned by rexsim_last_error)

constexpr int AG_MAX_O = 32, AG_MAX_A = 8, AG_MAX_H1 = 256, AG_MAX_H2 = 128;
constexpr int RED_BLOCKS = 256;    // partial blocks of the filter update

struct AgentDev {
    RexAgentConfig cfg;
    const float* __restrict__ params;    // policy block | value block
    int pol_floats, val_floats;
    float* filt;                          // observ mean[O], var_sum[O], reward mean, var_sum
    int32_t* cnt;                         // observ count, reward count, step counter, blocks-done ticket
    float* partial;                       // [RED_BLOCKS][2 * (O + 1)]
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// one 1-D TMA bulk copy global -> shared (weights of one network), completion on an mbarrier with explicit parity
__device__ __forceinline__ void tma_bulk_load(float* smem_dst, const float* gsrc, uint32_t bytes, uint64_t* bar, uint32_t parity) {
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t dst_a = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy reads of the buffer vs the async write
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_a), "l"(gsrc), "r"(bytes), "r"(bar_a) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar_a), "r"(parity) : "memory");
    }
}

struct PerformArgs {
    AgentDev D;
    const float* __restrict__ observ;
    int n, training;
    uint64_t seed;
    uint32_t step, env_offset;
    float* action; float* mean; float* logprob; float* value; float* observ_copy;
};

// Persistent CTAs, blockIdx.y = network (0 policy, 1 value): weights -> shared memory once (TMA bulk copy), then for every
// tile of TM envs:  normalise -> layer 1 (O x H1) -> layer 2 (H1 x H2: TM/16 envs x 4 outputs per thread, register tile) ->
// head (H2 x A | 1) from the register tile + a shared-memory reduction over the H2/4 column groups -> tanh / sampling.
// blockDim.x = 16 * (H2 / 4)  (400 threads for the reference's 200-100 networks).
template <int TM>
__global__ void __launch_bounds__(512, 1) perform_kernel(const PerformArgs P) {
    constexpr int RE = TM / 16;                           // envs per thread in the layer-2 register tile
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const RexAgentConfig& c = P.D.cfg;
    const int O = c.obs_dim, A = c.action_dim, H1 = c.hidden1, H2 = c.hidden2;
    const int nt = blockDim.x, tid = threadIdx.x;
    const int net = blockIdx.y;
    const int wmax = max(P.D.pol_floats, P.D.val_floats);
    float* W = smem;                          // weights of this CTA's network
    float* xs = W + wmax;                     // [O][TM] normalised observations
    float* h1 = xs + AG_MAX_O * TM;           // [H1][TM]; reused for the head partials [H2/4][TM][AO]
    float* nrm = h1 + (size_t)H1 * TM;        // [2][O] mean, 1/(std + 1e-8)
    if (tid == 0) {
        uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < O) {   // StreamingNormalize.transform constants (normalize.py:57-66, _std :131-144)
        const int cnt = P.D.cnt[0];
        float m = P.D.filt[tid], vs = P.D.filt[O + tid];
        float inv = 1.f;
        if (cnt > 1) inv = 1.f / (sqrtf(vs / (float)(cnt - 1) + 1e-4f) + 1e-8f);
        nrm[tid] = m; nrm[O + tid] = inv;
    }
    __syncthreads();
    const int ntiles = (P.n + TM - 1) / TM;
    const uint32_t step = P.step + (uint32_t)P.D.cnt[2];
    const int eg = tid & 15, og = tid >> 4;
    const int ngroups = H2 / 4;
    const bool l2_active = og < ngroups;              // the launch rounds the block up to TM threads for tiny networks
    const int AO = net == 0 ? A : 1;
    const int wfloats = net == 0 ? P.D.pol_floats : P.D.val_floats;
    tma_bulk_load(W, P.D.params + (net == 0 ? 0 : P.D.pol_floats), (uint32_t)wfloats * 4u, &bar, 0u);
    const float* W1 = W; const float* b1 = W1 + O * H1;
    const float* W2 = b1 + H1; const float* b2 = W2 + H1 * H2;
    const float* W3 = b2 + H2; const float* b3 = W3 + H2 * AO;
    const float* logstd = b3 + AO;                      // policy block only
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int e0 = tile * TM;
        __syncthreads();                                // previous tile's head is done with h1
        for (int i = tid; i < TM * O; i += nt) {        // coalesced read of the [TM][O] block
            const int e = i / O, o = i - e * O;
            float v = 0.f;
            if (e0 + e < P.n) {
                v = P.observ[(size_t)(e0 + e) * O + o];
                if (net == 0 && P.observ_copy) P.observ_copy[(size_t)(e0 + e) * O + o] = v;
            }
            v = (v - nrm[o]) * nrm[O + o];
            v = fminf(fmaxf(v, -c.observ_clip), c.observ_clip);
            xs[o * TM + e] = v;
        }
        __syncthreads();
        for (int i = tid; i < H1 * TM; i += nt) {       // layer 1
            const int k = i / TM, e = i - k * TM;
            float s = b1[k];
            for (int o = 0; o < O; o++) s = fmaf(xs[o * TM + e], W1[o * H1 + k], s);
            h1[i] = fmaxf(s, 0.f);
        }
        __syncthreads();
        // layer 2 register tile, held as env PAIRS so the inner product runs on Blackwell's packed fp32x2 FMA (FFMA2, sm_100):
        // acc2[p][j] = (env 2p, env 2p+1) x hidden unit j; per k one LDS.128 of weights, RE/4 LDS.128 of activations and
        // 2*RE FFMA2 (= 4*RE fp32 FMAs in half the issue slots; the weight is the instruction's scalar-broadcast operand)
        float2 acc2[RE / 2][4];
        if (l2_active) {
            const float4 bb = ld4(b2 + 4 * og);
#pragma unroll
            for (int p = 0; p < RE / 2; p++) {
                acc2[p][0] = make_float2(bb.x, bb.x); acc2[p][1] = make_float2(bb.y, bb.y);
                acc2[p][2] = make_float2(bb.z, bb.z); acc2[p][3] = make_float2(bb.w, bb.w);
            }
            // thread's envs: chunk c (of RE/4) covers envs 64*c + 4*eg .. +3, so the 16 lanes of a half-warp read 256
            // contiguous bytes per LDS.128 (bank-conflict free; envs 8*eg .. 8*eg+7 would be a 4-way conflict)
            const float* hp = h1 + 4 * eg;
            const float* wp = W2 + 4 * og;
#pragma unroll 4
            for (int k = 0; k < H1; k++) {
                const float4 w = ld4(wp);
                const float2 w0 = make_float2(w.x, w.x), w1 = make_float2(w.y, w.y), w2 = make_float2(w.z, w.z), w3 = make_float2(w.w, w.w);
#pragma unroll
                for (int cidx = 0; cidx < RE / 4; cidx++) {
                    const float4 t = ld4(hp + cidx * 64);
                    const float2 h01 = make_float2(t.x, t.y), h23 = make_float2(t.z, t.w);
                    acc2[2 * cidx][0] = __ffma2_rn(h01, w0, acc2[2 * cidx][0]); acc2[2 * cidx][1] = __ffma2_rn(h01, w1, acc2[2 * cidx][1]);
                    acc2[2 * cidx][2] = __ffma2_rn(h01, w2, acc2[2 * cidx][2]); acc2[2 * cidx][3] = __ffma2_rn(h01, w3, acc2[2 * cidx][3]);
                    acc2[2 * cidx + 1][0] = __ffma2_rn(h23, w0, acc2[2 * cidx + 1][0]); acc2[2 * cidx + 1][1] = __ffma2_rn(h23, w1, acc2[2 * cidx + 1][1]);
                    acc2[2 * cidx + 1][2] = __ffma2_rn(h23, w2, acc2[2 * cidx + 1][2]); acc2[2 * cidx + 1][3] = __ffma2_rn(h23, w3, acc2[2 * cidx + 1][3]);
                }
                hp += TM; wp += H2;
            }
        }
        float acc[RE][4];                               // unpack: acc[i][j], i = 4*chunk + lane-in-chunk
#pragma unroll
        for (int p = 0; p < RE / 2; p++)
#pragma unroll
            for (int j = 0; j < 4; j++) { acc[2 * p][j] = acc2[p][j].x; acc[2 * p + 1][j] = acc2[p][j].y; }
        __syncthreads();                                // all reads of h1 done: reuse it for the head partials
        if (l2_active) {                                // head, part 1: this thread's 4 hidden units x its RE envs
            float w3[4][AG_MAX_A];
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int a = 0; a < AG_MAX_A; a++) w3[j][a] = a < AO ? W3[(4 * og + j) * AO + a] : 0.f;
#pragma unroll
            for (int i = 0; i < RE; i++) {
                const float r0 = fmaxf(acc[i][0], 0.f), r1 = fmaxf(acc[i][1], 0.f), r2 = fmaxf(acc[i][2], 0.f), r3 = fmaxf(acc[i][3], 0.f);
                float* dst = h1 + ((size_t)og * TM + (i / 4) * 64 + 4 * eg + (i & 3)) * AO;
#pragma unroll
                for (int a = 0; a < AG_MAX_A; a++)
                    if (a < AO) dst[a] = fmaf(r0, w3[0][a], fmaf(r1, w3[1][a], fmaf(r2, w3[2][a], r3 * w3[3][a])));
            }
        }
        __syncthreads();
        if (tid < TM && e0 + tid < P.n) {               // head, part 2: one thread per env sums the column groups in order
            const int e = tid, env = e0 + e;
            float s[AG_MAX_A];
#pragma unroll
            for (int a = 0; a < AG_MAX_A; a++) s[a] = a < AO ? b3[a] : 0.f;
            for (int g = 0; g < ngroups; g++) {
                const float* src = h1 + ((size_t)g * TM + e) * AO;
#pragma unroll
                for (int a = 0; a < AG_MAX_A; a++) if (a < AO) s[a] += src[a];
            }
            if (net == 1) {
                if (P.value) P.value[env] = s[0];
            } else {
                float lp = 0.f;
#pragma unroll
                for (int a = 0; a < AG_MAX_A; a++) {
                    if (a < A) {
                        const float mu = tanhf(s[a]);
                        const float ls = logstd[a];
                        float act = mu, z = 0.f;
                        if (P.training) {               // network.policy.sample (algorithm.py:116)
                            const uint32_t genv = P.env_offset + (uint32_t)env;
                            const float u1 = ((float)(rand_u32(P.seed, genv, step, 2 * a) >> 8) + 0.5f) * (1.0f / 16777216.0f);
                            const float u2 = ((float)(rand_u32(P.seed, genv, step, 2 * a + 1) >> 8) + 0.5f) * (1.0f / 16777216.0f);
                            z = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
                            act = fmaf(expf(ls), z, mu);
                        }
                        lp += -0.5f * z * z - ls - 0.9189385332046727f;     // diag-normal log-density
                        if (P.action) P.action[(size_t)env * A + a] = act;
                        if (P.mean) P.mean[(size_t)env * A + a] = mu;
                    }
                }
                if (P.logprob) P.logprob[env] = lp;
            }
        }
    }
}


// -------------------------------------------------------------------------------------------------------------------
// perform on the 5th-generation tensor cores (opt-in: rexagent_set_precision(a, 1)).
//
// Layer 2 (H1 x H2 = 200 x 100, 97 % of the network's flops) runs as ONE tcgen05.mma chain per 128-env tile:
//   D[128 envs][N] (fp32, TMEM) = A[128][H1] (layer-1 activations, shared memory) x B[N][H1]^T (W2^T, shared memory),
// kind::tf32 (operands rounded to 10-bit mantissas with cvt.rna, fp32 accumulation), M = 128, N = H2 rounded up to 16,
// K = 8 per instruction, H1 / 8 instructions issued by one thread; completion arrives on an mbarrier (tcgen05.commit);
// the eight warps read the accumulators back (tcgen05.ld 32x32b: a warp reaches the TMEM lanes of its quarter, 32 envs; two
// warps per quarter split the columns) and finish bias, ReLU, head and sampling one env per thread.  Layer 1 (O <= 16 inputs) and the head stay on the CUDA cores in fp32.
//
// Operand layout in shared memory: K-major, no swizzle (UMMA "interleave"): 8-row x 16-byte core matrices (8 rows x 4
// tf32), the two core matrices one instruction reads along K are LBO bytes apart, 8-row groups SBO = 128 bytes apart:
//   byte(row, k) = (k / 4) * LBO + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4,   LBO = rows * 16.
// Layer 1 writes A directly in that form (one 16-byte store per (env, 4 hidden units)).
// TF32 keeps 11 significant bits: the measured error against the fp32 kernel is stated in DESIGN.md; the fp32 FFMA2
// kernel above stays the parity default.
// -------------------------------------------------------------------------------------------------------------------
namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets (16-byte units),
// version 1 (Blackwell), no swizzle, base offset 0
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_a, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar_a), "r"(parity) : "memory");
    }
}
constexpr int TM = 128, THREADS = 256, TMEM_COLS = 128;
}  // namespace tc

static size_t perform_tc_smem_bytes(const RexAgentConfig* c) {
    const int npad = (c->hidden2 + 15) & ~15;
    const size_t a = (size_t)tc::TM * c->hidden1 * 4, b = (size_t)npad * c->hidden1 * 4;
    const size_t small = ((size_t)c->obs_dim * c->hidden1 + c->hidden1 + c->hidden2 + (size_t)c->hidden2 * AG_MAX_A + 2 * AG_MAX_A +
                          (size_t)c->obs_dim * tc::TM + 2 * AG_MAX_O) * 4;
    return a + b + small + 64;
}

// OT / AT: obs_dim / action_dim known at compile time (0 = generic: loops over the maxima with predicates -- every predicated-off
// instruction still takes an issue slot, which made the generic form 4x heavier than needed for the reference's 4-in / 2-out nets)
template <int OT, int AT>
__global__ void __launch_bounds__(tc::THREADS, 1) perform_tc_kernel(const PerformArgs P) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smraw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const RexAgentConfig& c = P.D.cfg;
    const int O = OT ? OT : c.obs_dim, A = AT ? AT : c.action_dim, H1 = c.hidden1, H2 = c.hidden2;
    constexpr int OU = OT ? OT : 16;                        // unroll bound of the input loops
    const int NP = (H2 + 15) & ~15;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int net = blockIdx.y;
    const int AO = net == 0 ? A : 1;
    const uint32_t LBO_A = TM * 16, LBO_B = (uint32_t)NP * 16;
    uint8_t* Asm = smraw;                                   // [H1/4][TM/8][8][16 B]
    uint8_t* Bsm = Asm + (size_t)TM * H1 * 4;               // [H1/4][NP/8][8][16 B]
    float* W1 = reinterpret_cast<float*>(Bsm + (size_t)NP * H1 * 4);   // [O][H1]
    float* b1 = W1 + O * H1;
    float* b2 = b1 + H1;
    float* W3 = b2 + H2;                                    // [H2][AO]
    float* b3 = W3 + H2 * AG_MAX_A;                         // b3[AO], then logstd[AO]
    float* xs = b3 + 2 * AG_MAX_A;                          // [O][TM]
    float* nrm = xs + O * TM;                               // [2][O]
    const float* G = P.D.params + (net == 0 ? 0 : P.D.pol_floats);
    const float* gW1 = G; const float* gb1 = gW1 + O * H1; const float* gW2 = gb1 + H1; const float* gb2 = gW2 + H1 * H2;
    const float* gW3 = gb2 + H2; const float* gb3 = gW3 + H2 * AO;
    const uint32_t bar_a = smem_u32(&bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                        // one warp allocates the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < O) {   // StreamingNormalize.transform constants (normalize.py:57-66, _std :131-144)
        const int cnt = P.D.cnt[0];
        float m = P.D.filt[tid], vs = P.D.filt[O + tid];
        float inv = 1.f;
        if (cnt > 1) inv = 1.f / (sqrtf(vs / (float)(cnt - 1) + 1e-4f) + 1e-8f);
        nrm[tid] = m; nrm[O + tid] = inv;
    }
    // weights: the small layers as they are, W2 transposed into the B operand form (rounded to tf32, zero rows up to NP)
    for (int i = tid; i < O * H1; i += THREADS) W1[i] = gW1[i];
    for (int i = tid; i < H1; i += THREADS) b1[i] = gb1[i];
    for (int i = tid; i < H2; i += THREADS) b2[i] = gb2[i];
    for (int i = tid; i < H2 * AO; i += THREADS) W3[i] = gW3[i];
    if (tid < 2 * AO) b3[tid] = gb3[tid];                   // b3[AO] followed by logstd[AO] (policy block; unused for the value net)
    {   // W2 [H1][H2] row-major -> B[n][k]: 16-byte global loads (4 consecutive n of one k), several in flight per thread
        const int nq = NP >> 2;
#pragma unroll 4
        for (int i = tid; i < nq * H1; i += THREADS) {
            const int k = i / nq, n = (i - k * nq) << 2;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < H2) w = *reinterpret_cast<const float4*>(gW2 + (size_t)k * H2 + n);        // H2 is a multiple of 4
            uint8_t* dst = Bsm + (size_t)(k >> 2) * LBO_B + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4;
            *reinterpret_cast<float*>(dst) = to_tf32(w.x); *reinterpret_cast<float*>(dst + 16) = to_tf32(w.y);
            *reinterpret_cast<float*>(dst + 32) = to_tf32(w.z); *reinterpret_cast<float*>(dst + 48) = to_tf32(w.w);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // B was written through the generic proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = instr_desc_tf32(TM, NP);
    const uint32_t a_addr = smem_u32(Asm), b_addr = smem_u32(Bsm);
    const int ntiles = (P.n + TM - 1) / TM;
    const uint32_t step = P.step + (uint32_t)P.D.cnt[2];
    uint32_t phase = 0;
    // normalised observations of a tile -> xs (coalesced read of the [TM][O] block).  Issued for the first tile here and for
    // tile t + 1 right after layer 1 of tile t, so the global loads are in flight while the MMA chain and the epilogue run.
    auto load_obs = [&](int tile) {
        const int e0 = tile * TM;
        for (int i = tid; i < TM * O; i += THREADS) {
            const int e = i / O, o = i - e * O;
            float v = 0.f;
            if (e0 + e < P.n) {
                v = P.observ[(size_t)(e0 + e) * O + o];
                if (net == 0 && P.observ_copy) P.observ_copy[(size_t)(e0 + e) * O + o] = v;
            }
            v = (v - nrm[o]) * nrm[O + o];
            v = fminf(fmaxf(v, -c.observ_clip), c.observ_clip);
            xs[o * TM + e] = v;
        }
    };
    if ((int)blockIdx.x < ntiles) load_obs(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, phase ^= 1u) {
        const int e0 = tile * TM;
        __syncthreads();
        // layer 1 -> A operand: thread = (env, half of the hidden units); the env's inputs stay in registers, the weight rows
        // are warp-wide broadcast loads, every 4 hidden units leave as one 16-byte store into the core-matrix layout
        {
            const int e = tid & (TM - 1), half = tid >> 7;              // THREADS == 2 * TM
            float x[OU];
#pragma unroll
            for (int o = 0; o < OU; o++) x[o] = (OT || o < O) ? xs[o * TM + e] : 0.f;
            const int nk4 = H1 >> 2, k_lo = half ? (nk4 + 1) / 2 : 0, k_hi = half ? nk4 : (nk4 + 1) / 2;
            uint8_t* dst = Asm + (e >> 3) * 128 + (e & 7) * 16;
#pragma unroll 2
            for (int k4 = k_lo; k4 < k_hi; k4++) {
                float4 s4 = *reinterpret_cast<const float4*>(b1 + 4 * k4);
#pragma unroll
                for (int o = 0; o < OU; o++) {
                    if (OT || o < O) {
                        const float4 w = *reinterpret_cast<const float4*>(W1 + o * H1 + 4 * k4);
                        s4.x = fmaf(x[o], w.x, s4.x); s4.y = fmaf(x[o], w.y, s4.y); s4.z = fmaf(x[o], w.z, s4.z); s4.w = fmaf(x[o], w.w, s4.w);
                    }
                }
                *reinterpret_cast<float4*>(dst + (size_t)k4 * LBO_A) =
                    make_float4(to_tf32(fmaxf(s4.x, 0.f)), to_tf32(fmaxf(s4.y, 0.f)), to_tf32(fmaxf(s4.z, 0.f)), to_tf32(fmaxf(s4.w, 0.f)));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // A: generic-proxy writes -> tensor-core (async proxy) reads
        __syncthreads();
        if (tid == 0) {                                     // the MMA chain: H1 / 8 instructions, one issuing thread
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int kb = 0; kb < (H1 >> 3); kb++) {
                const uint64_t da = smem_desc(a_addr + (uint32_t)kb * 2u * LBO_A, LBO_A, 128u);
                const uint64_t db = smem_desc(b_addr + (uint32_t)kb * 2u * LBO_B, LBO_B, 128u);
                mma_tf32(tmem_base, da, db, idesc, kb > 0 ? 1u : 0u);
            }
            // completion of everything issued so far -> one arrival on the mbarrier (implies tcgen05.fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_a) : "memory");
        }
        if (tile + (int)gridDim.x < ntiles) load_obs(tile + gridDim.x);      // xs is free again: next tile's loads fly during the MMA chain
        auto epilogue = [&](auto ao_tag) {
            constexpr int AC = decltype(ao_tag)::value;          // head width at compile time (0: generic, predicated up to AG_MAX_A)
            constexpr int AU = AC ? AC : AG_MAX_A;
            // epilogue on all 8 warps: warp w may read TMEM lanes 32 (w % 4) .. +31, so warps w and w + 4 share a quarter of the
            // envs and split the accumulator columns in two halves; thread <-> env e0 + 32 (w % 4) + lane.  The upper half hands
            // its partial head sums over through shared memory (the A operand's first bytes: the MMA chain is done with it).
            mbar_wait(bar_a, phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int q = warp & 3, hi = warp >> 2;
            const int el = q * 32 + (tid & 31);
            const int env = e0 + el;
            float sh[AU];
#pragma unroll
            for (int a = 0; a < AU; a++) sh[a] = ((AC || a < AO) && !hi) ? b3[a] : 0.f;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
            const int chunks = NP >> 4, c_lo = hi ? (chunks + 1) / 2 : 0, c_hi = hi ? chunks : (chunks + 1) / 2;
            for (int cc = c_lo; cc < c_hi; cc++) {
                const int c0 = cc << 4;
                uint32_t r[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(lane_addr + (uint32_t)c0) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                    const int n = c0 + 4 * j4;
                    if (n < H2) {                                       // H2 is a multiple of 4: whole groups only
                        const float4 bb = *reinterpret_cast<const float4*>(b2 + n);
                        const float h0 = fmaxf(__uint_as_float(r[4 * j4]) + bb.x, 0.f), h1v = fmaxf(__uint_as_float(r[4 * j4 + 1]) + bb.y, 0.f);
                        const float h2v = fmaxf(__uint_as_float(r[4 * j4 + 2]) + bb.z, 0.f), h3v = fmaxf(__uint_as_float(r[4 * j4 + 3]) + bb.w, 0.f);
                        const float* w3 = W3 + n * AO;
#pragma unroll
                        for (int a = 0; a < AU; a++)
                            if (AC || a < AO) sh[a] = fmaf(h3v, w3[3 * AO + a], fmaf(h2v, w3[2 * AO + a], fmaf(h1v, w3[AO + a], fmaf(h0, w3[a], sh[a]))));
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            float* part = reinterpret_cast<float*>(Asm);    // [AG_MAX_A][TM] partial sums of the upper column half
            if (hi) {
#pragma unroll
                for (int a = 0; a < AU; a++) if (AC || a < AO) part[a * TM + el] = sh[a];
            }
            __syncthreads();
            if (!hi) {
#pragma unroll
                for (int a = 0; a < AU; a++) if (AC || a < AO) sh[a] += part[a * TM + el];
            }
            if (!hi && env < P.n) {
                if (net == 1) {
                    if (P.value) P.value[env] = sh[0];
                } else {
                    float lp = 0.f;
#pragma unroll
                    for (int a = 0; a < AU; a++) {
                        if (AC || a < A) {
                            const float mu = tanhf(sh[a]);
                            const float ls = b3[AO + a];
                            float act = mu, z = 0.f;
                            if (P.training) {               // network.policy.sample (algorithm.py:116)
                                const uint32_t genv = P.env_offset + (uint32_t)env;
                                const float u1 = ((float)(rand_u32(P.seed, genv, step, 2 * a) >> 8) + 0.5f) * (1.0f / 16777216.0f);
                                const float u2 = ((float)(rand_u32(P.seed, genv, step, 2 * a + 1) >> 8) + 0.5f) * (1.0f / 16777216.0f);
                                z = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
                                act = fmaf(expf(ls), z, mu);
                            }
                            lp += -0.5f * z * z - ls - 0.9189385332046727f;     // diag-normal log-density
                            if (P.action) P.action[(size_t)env * A + a] = act;
                            if (P.mean) P.mean[(size_t)env * A + a] = mu;
                        }
                    }
                    if (P.logprob) P.logprob[env] = lp;
                }
            }
        };
        if (net == 1) epilogue(std::integral_constant<int, 1>{});
        else epilogue(std::integral_constant<int, AT>{});
        __syncthreads();                                    // accumulators read, A free: next tile
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// StreamingNormalize.update for the observation columns and the reward column in one launch (normalize.py:73-99):
// per column j: S1 = sum(x - mean), S2 = sum((x - mean)^2) over the batch (block partials, then the last block to
// finish adds them in a fixed order -> bit-reproducible); count += n; new_mean = mean + S1/count;
// var_sum += S2 - S1^2/count  ( = sum (x - mean)(x - new_mean) ).
struct ExperienceArgs { AgentDev D; const float* __restrict__ observ; const float* __restrict__ reward; int n; float* sums_out; };

// finalisation shared by the single-GPU kernel and the sharded path (sums = [O+1][2] (S1, S2) over ALL ranks, n_total likewise)
__device__ __forceinline__ void filters_finalize(const AgentDev& D, int col, float s1, float s2, int n_total, float first_value) {
    const int O = D.cfg.obs_dim;
    const int which = col < O ? 0 : 1;
    const int cnt = D.cnt[which] + n_total;
    const float step = (float)cnt;
    float* meanp = col < O ? &D.filt[col] : &D.filt[2 * O];
    float* varp = col < O ? &D.filt[O + col] : &D.filt[2 * O + 1];
    float new_mean = *meanp + s1 / step;
    if (cnt <= 1) new_mean = first_value;                                   // tf.cond(count > 1, new_mean, value[0])
    *varp += s2 - s1 * (s1 / step);
    *meanp = new_mean;
}

__global__ void __launch_bounds__(256) experience_kernel(const ExperienceArgs P) {
    const int O = P.D.cfg.obs_dim, C = O + 1;
    __shared__ float red[2 * (AG_MAX_O + 1)][8];
    __shared__ float mean_s[AG_MAX_O + 1];
    __shared__ int last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < C) mean_s[tid] = tid < O ? P.D.filt[tid] : P.D.filt[2 * O];
    __syncthreads();
    // each thread walks rows with a fixed stride; column sums kept per thread for its own column set
    for (int cidx = 0; cidx < C; cidx++) {
        float s1 = 0.f, s2 = 0.f;
        const float m = mean_s[cidx];
        for (int e = blockIdx.x * blockDim.x + tid; e < P.n; e += gridDim.x * blockDim.x) {
            const float x = cidx < O ? P.observ[(size_t)e * O + cidx] : P.reward[e];
            const float d = x - m;
            s1 += d; s2 = fmaf(d, d, s2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (lane == 0) { red[2 * cidx][warp] = s1; red[2 * cidx + 1][warp] = s2; }
    }
    __syncthreads();
    if (tid < 2 * C) {
        float s = 0.f;
        for (int w = 0; w < 8; w++) s += red[tid][w];
        P.D.partial[(size_t)blockIdx.x * 2 * C + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = (atomicAdd(&P.D.cnt[3], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (tid < C) {
        float s1 = 0.f, s2 = 0.f;
        for (int b = 0; b < (int)gridDim.x; b++) {
            s1 += __ldcg(&P.D.partial[(size_t)b * 2 * C + 2 * tid]);
            s2 += __ldcg(&P.D.partial[(size_t)b * 2 * C + 2 * tid + 1]);
        }
        if (P.sums_out) { P.sums_out[2 * tid] = s1; P.sums_out[2 * tid + 1] = s2; }          // sharded: another rank's sums are added first
        else filters_finalize(P.D, tid, s1, s2, P.n, tid < O ? P.observ[tid] : P.reward[0]);
    }
    __syncthreads();
    if (tid == 0) {
        if (!P.sums_out) { P.D.cnt[0] += P.n; P.D.cnt[1] += P.n; P.D.cnt[2] += 1; }     // step counter: the next perform draws fresh noise
        P.D.cnt[3] = 0;
    }
}

// sharded rollouts: sums = this rank's partial sums after an all-reduce(sum) over the ranks; n_total = envs of all ranks
__global__ void experience_finalize_kernel(const AgentDev D, const float* __restrict__ sums, int n_total,
                                           const float* __restrict__ observ, const float* __restrict__ reward) {
    const int O = D.cfg.obs_dim, tid = threadIdx.x;
    if (tid <= O) filters_finalize(D, tid, sums[2 * tid], sums[2 * tid + 1], n_total, tid < O ? observ[tid] : reward[0]);
    __syncthreads();
    if (tid == 0) { D.cnt[0] += n_total; D.cnt[1] += n_total; D.cnt[2] += 1; }
}

__global__ void transform_reward_kernel(const AgentDev D, const float* __restrict__ r, int n, float* __restrict__ out) {
    const int O = D.cfg.obs_dim;
    const int cnt = D.cnt[1];
    float inv = 1.f;
    if (cnt > 1) inv = 1.f / (sqrtf(D.filt[2 * O + 1] / (float)(cnt - 1) + 1e-4f) + 1e-8f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = fminf(fmaxf(r[i] * inv, -D.cfg.reward_clip), D.cfg.reward_clip);      // center=False, scale=True, clip=10
}

// ---- scans ---------------------------------------------------------------------------------------------------------
// one thread per episode row, walking backwards in time; rows are laid out by (stride_e, stride_t) so a time-major
// buffer (stride_e = 1) is read fully coalesced
// The recurrences are serial in time but every load is independent of them: each thread fetches SCAN_U time steps into
// registers first (SCAN_U x 2-3 loads in flight per thread), then runs the dependent chain -- without this the kernel waits
// one DRAM latency per time step.
constexpr int SCAN_U = 8;

template <bool ADV>
__global__ void __launch_bounds__(256) row_scan_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                                                       const int32_t* __restrict__ length, int episodes, int L,
                                                       int64_t se, int64_t st, float discount, float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= episodes) return;
    const int len = length[e];
    float agg = 0.f, next_v = 0.f;                           // value[:, 1:] padded with zeros (utility.py:116)
    const int64_t base = (int64_t)e * se;
    for (int t1 = L; t1 > 0; t1 -= SCAN_U) {
        float r[SCAN_U], v[SCAN_U];
#pragma unroll
        for (int u = 0; u < SCAN_U; u++) {
            const int t = t1 - 1 - u;
            r[u] = t >= 0 ? reward[base + (int64_t)t * st] : 0.f;
            v[u] = (ADV && t >= 0) ? value[base + (int64_t)t * st] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < SCAN_U; u++) {
            const int t = t1 - 1 - u;
            if (t < 0) break;
            float cur;
            if (ADV) { cur = (t < len) ? r[u] + discount * next_v - v[u] : 0.f; next_v = v[u]; }   // mask * delta
            else cur = (t < len) ? r[u] : 0.f;                                                     // mask * reward
            agg = cur + discount * agg;
            out[base + (int64_t)t * st] = agg;
        }
    }
}

__global__ void __launch_bounds__(256) gae_segments_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                                                           const uint8_t* __restrict__ done, int T, int n, float discount,
                                                           float lambda, float* __restrict__ out_ret, float* __restrict__ out_adv) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float next_v = value[(size_t)T * n + e], ret = next_v, adv = 0.f;
    for (int t1 = T; t1 > 0; t1 -= SCAN_U) {
        float r[SCAN_U], v[SCAN_U]; uint8_t d[SCAN_U];
#pragma unroll
        for (int u = 0; u < SCAN_U; u++) {
            const int t = t1 - 1 - u;
            const size_t idx = (size_t)(t >= 0 ? t : 0) * n + e;
            r[u] = reward[idx]; v[u] = value[idx]; d[u] = done[idx];
        }
#pragma unroll
        for (int u = 0; u < SCAN_U; u++) {
            const int t = t1 - 1 - u;
            if (t < 0) break;
            const size_t idx = (size_t)t * n + e;
            const float nd = d[u] ? 0.f : 1.f;
            const float delta = r[u] + discount * next_v * nd - v[u];
            adv = delta + discount * lambda * nd * adv;
            ret = r[u] + discount * nd * ret;
            if (out_adv) out_adv[idx] = adv;
            if (out_ret) out_ret[idx] = ret;
            next_v = v[u];
        }
    }
}

}  // namespace rexsim

using namespace rexsim;

struct RexAgent {
    AgentDev D;
    float* d_params = nullptr;
    float* d_filt = nullptr;
    int32_t* d_cnt = nullptr;
    float* d_partial = nullptr;
    int sm_count = 148;
    int64_t launches = 0;
    int precision = 0;                 // 0: fp32 on the CUDA cores (parity default); 1: layer 2 on the tensor cores in TF32
};

// the instance compiled for this network's input / output widths (the reference's tasks: 4 or 16 observations, 1-8 actions)
typedef void (*PerformTcKernel)(const PerformArgs);
static PerformTcKernel perform_tc_pick(const RexAgentConfig* c) {
    const int O = c->obs_dim, A = c->action_dim;
    if (O == 4) { if (A == 1) return perform_tc_kernel<4, 1>; if (A == 2) return perform_tc_kernel<4, 2>; if (A == 8) return perform_tc_kernel<4, 8>; return perform_tc_kernel<4, 0>; }
    if (O == 16) { if (A == 2) return perform_tc_kernel<16, 2>; if (A == 4) return perform_tc_kernel<16, 4>; return perform_tc_kernel<16, 0>; }
    return perform_tc_kernel<0, 0>;
}

static int afail(int code, const char* msg) { set_error(msg); return code; }
#define ACK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { char b[256]; snprintf(b, sizeof(b), "%s: %s", #x, cudaGetErrorString(_e)); set_error(b); return REXSIM_ERR_CUDA; } } while (0)

static int64_t pad4(int64_t n) { return (n + 3) & ~(int64_t)3; }
static int check_cfg(const RexAgentConfig* c) {
    if (!c) return afail(REXSIM_ERR_INVALID, "null agent config");
    if (c->obs_dim <= 0 || c->obs_dim > AG_MAX_O || c->action_dim <= 0 || c->action_dim > AG_MAX_A)
        return afail(REXSIM_ERR_INVALID, "agent: obs_dim must be 1..32 and action_dim 1..8");
    if (c->hidden1 <= 0 || c->hidden1 > AG_MAX_H1 || c->hidden2 <= 0 || c->hidden2 > AG_MAX_H2 || (c->hidden1 & 3) || (c->hidden2 & 3))
        return afail(REXSIM_ERR_UNSUPPORTED, "agent: hidden1 <= 256, hidden2 <= 128, both multiples of 4");
    return REXSIM_OK;
}

extern "C" {

int64_t rexagent_policy_floats(const RexAgentConfig* c) {
    if (!c) return 0;
    const int64_t O = c->obs_dim, A = c->action_dim, H1 = c->hidden1, H2 = c->hidden2;
    return pad4(O * H1 + H1 + H1 * H2 + H2 + H2 * A + A + A);
}
int64_t rexagent_value_floats(const RexAgentConfig* c) {
    if (!c) return 0;
    const int64_t O = c->obs_dim, H1 = c->hidden1, H2 = c->hidden2;
    return pad4(O * H1 + H1 + H1 * H2 + H2 + H2 + 1);
}

static size_t perform_smem_bytes(const RexAgentConfig* c, int tm) {
    const size_t wmax = (size_t)(rexagent_policy_floats(c) > rexagent_value_floats(c) ? rexagent_policy_floats(c) : rexagent_value_floats(c));
    return (wmax + (size_t)AG_MAX_O * tm + (size_t)c->hidden1 * tm + 2 * AG_MAX_O) * sizeof(float);
}

int rexagent_create(const RexAgentConfig* cfg, RexAgent** out) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!out) return afail(REXSIM_ERR_INVALID, "null argument");
    if (perform_smem_bytes(cfg, 128) > 227 * 1024) return afail(REXSIM_ERR_UNSUPPORTED, "agent: network does not fit the 227 KB of shared memory");
    if (cfg->action_dim * (cfg->hidden2 / 4) > cfg->hidden1) return afail(REXSIM_ERR_UNSUPPORTED, "agent: action_dim * hidden2 / 4 must not exceed hidden1");
    RexAgent* a = new RexAgent();
    struct Guard { RexAgent* a; ~Guard() { if (a) rexagent_destroy(a); } } guard{a};  // frees everything on an early return
    a->D.cfg = *cfg;
    a->D.pol_floats = (int)rexagent_policy_floats(cfg); a->D.val_floats = (int)rexagent_value_floats(cfg);
    const size_t np = (size_t)a->D.pol_floats + a->D.val_floats;
    const int O = cfg->obs_dim;
    int dev = 0;
    ACK(cudaGetDevice(&dev));
    ACK(cudaDeviceGetAttribute(&a->sm_count, cudaDevAttrMultiProcessorCount, dev));
    ACK(cudaMalloc(&a->d_params, np * sizeof(float)));
    ACK(cudaMemset(a->d_params, 0, np * sizeof(float)));
    ACK(cudaMalloc(&a->d_filt, (2 * O + 2) * sizeof(float)));
    ACK(cudaMemset(a->d_filt, 0, (2 * O + 2) * sizeof(float)));
    ACK(cudaMalloc(&a->d_cnt, 4 * sizeof(int32_t)));
    ACK(cudaMemset(a->d_cnt, 0, 4 * sizeof(int32_t)));
    ACK(cudaMalloc(&a->d_partial, (size_t)RED_BLOCKS * 2 * (O + 1) * sizeof(float)));
    ACK(cudaFuncSetAttribute(perform_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)perform_smem_bytes(cfg, 64)));
    ACK(cudaFuncSetAttribute(perform_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)perform_smem_bytes(cfg, 128)));
    a->D.params = a->d_params; a->D.filt = a->d_filt; a->D.cnt = a->d_cnt; a->D.partial = a->d_partial;
    ACK(cudaDeviceSynchronize());
    guard.a = nullptr;
    *out = a;
    return REXSIM_OK;
}
void rexagent_destroy(RexAgent* a) {
    if (!a) return;
    cudaFree(a->d_params); cudaFree(a->d_filt); cudaFree(a->d_cnt); cudaFree(a->d_partial);
    delete a;
}
int rexagent_set_precision(RexAgent* a, int32_t mode) {
    if (!a) return afail(REXSIM_ERR_INVALID, "null argument");
    if (mode != 0 && mode != 1) return afail(REXSIM_ERR_INVALID, "agent: precision must be 0 (fp32) or 1 (tf32 tensor cores)");
    if (mode == 1) {
        const RexAgentConfig* c = &a->D.cfg;
        if ((c->hidden1 & 7) || c->obs_dim > 16 || perform_tc_smem_bytes(c) > 227 * 1024)
            return afail(REXSIM_ERR_UNSUPPORTED, "agent: the tensor-core path needs hidden1 % 8 == 0, obs_dim <= 16 and operands that fit 227 KB of shared memory");
        ACK(cudaFuncSetAttribute(perform_tc_pick(c), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)perform_tc_smem_bytes(c)));
    }
    a->precision = mode;
    return REXSIM_OK;
}
int rexagent_set_params(RexAgent* a, const float* h, int64_t n) {
    if (!a || !h) return afail(REXSIM_ERR_INVALID, "null argument");
    if (n != (int64_t)a->D.pol_floats + a->D.val_floats) return afail(REXSIM_ERR_INVALID, "agent: parameter block size mismatch");
    ACK(cudaMemcpy(a->d_params, h, n * sizeof(float), cudaMemcpyHostToDevice));
    return REXSIM_OK;
}
int rexagent_get_params(RexAgent* a, float* h, int64_t n) {
    if (!a || !h) return afail(REXSIM_ERR_INVALID, "null argument");
    if (n != (int64_t)a->D.pol_floats + a->D.val_floats) return afail(REXSIM_ERR_INVALID, "agent: parameter block size mismatch");
    ACK(cudaMemcpy(h, a->d_params, n * sizeof(float), cudaMemcpyDeviceToHost));
    return REXSIM_OK;
}
int rexagent_params_buffer(RexAgent* a, float** p) {
    if (!a || !p) return afail(REXSIM_ERR_INVALID, "null argument");
    *p = a->d_params;
    return REXSIM_OK;
}
int rexagent_state_buffers(RexAgent* a, float** filt, int32_t** counters) {
    if (!a) return afail(REXSIM_ERR_INVALID, "null argument");
    if (filt) *filt = a->d_filt;
    if (counters) *counters = a->d_cnt;
    return REXSIM_OK;
}
int rexagent_set_filters(RexAgent* a, int32_t oc, const float* om, const float* ov, int32_t rc, float rm, float rv) {
    if (!a || !om || !ov) return afail(REXSIM_ERR_INVALID, "null argument");
    const int O = a->D.cfg.obs_dim;
    std::vector<float> f(2 * O + 2);
    for (int i = 0; i < O; i++) { f[i] = om[i]; f[O + i] = ov[i]; }
    f[2 * O] = rm; f[2 * O + 1] = rv;
    int32_t c[4] = {oc, rc, 0, 0};
    ACK(cudaMemcpy(a->d_filt, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
    ACK(cudaMemcpy(a->d_cnt, c, sizeof(c), cudaMemcpyHostToDevice));
    return REXSIM_OK;
}
int rexagent_get_filters(RexAgent* a, int32_t* counts, float* om, float* ov, float* rmv) {
    if (!a || !counts || !om || !ov || !rmv) return afail(REXSIM_ERR_INVALID, "null argument");
    const int O = a->D.cfg.obs_dim;
    std::vector<float> f(2 * O + 2);
    int32_t c[4];
    ACK(cudaMemcpy(f.data(), a->d_filt, f.size() * sizeof(float), cudaMemcpyDeviceToHost));
    ACK(cudaMemcpy(c, a->d_cnt, sizeof(c), cudaMemcpyDeviceToHost));
    for (int i = 0; i < O; i++) { om[i] = f[i]; ov[i] = f[O + i]; }
    rmv[0] = f[2 * O]; rmv[1] = f[2 * O + 1];
    counts[0] = c[0]; counts[1] = c[1];
    return REXSIM_OK;
}

int rexagent_perform(RexAgent* a, const float* observ, int32_t n, int32_t training, uint64_t seed, uint32_t step,
                     uint32_t env_offset, float* action, float* mean, float* logprob, float* value, float* observ_copy,
                     void* stream) {
    if (!a || !observ) return afail(REXSIM_ERR_INVALID, "null argument");
    if (n <= 0) return afail(REXSIM_ERR_INVALID, "agent: n must be positive");
    PerformArgs P;
    P.D = a->D; P.observ = observ; P.n = n; P.training = training; P.seed = seed; P.step = step; P.env_offset = env_offset;
    P.action = action; P.mean = mean; P.logprob = logprob; P.value = value; P.observ_copy = observ_copy;
    if (a->precision == 1) {          // tcgen05 path: 128-env tiles, one CTA per SM, networks in different CTAs (blockIdx.y)
        const int half_tc = a->sm_count / 2 > 0 ? a->sm_count / 2 : 1;
        const int nt = (n + tc::TM - 1) / tc::TM;
        dim3 grid_tc(nt < half_tc ? nt : half_tc, 2);
        perform_tc_pick(&a->D.cfg)<<<grid_tc, tc::THREADS, perform_tc_smem_bytes(&a->D.cfg), (cudaStream_t)stream>>>(P);
        ACK(cudaGetLastError());
        a->launches++;
        return REXSIM_OK;
    }
    // one CTA per SM (the weights fill most of the shared memory); the two networks run in different CTAs (blockIdx.y).
    // Small batches use 64-env tiles so that more SMs get a tile, large ones 128-env tiles (8 x 4 register tile per thread).
    const int half = a->sm_count / 2 > 0 ? a->sm_count / 2 : 1;
    const bool big = (n + 127) / 128 >= half;
    const int tm = big ? 128 : 64;
    const int ntiles = (n + tm - 1) / tm;
    dim3 grid(ntiles < half ? ntiles : half, 2);
    int threads = 16 * (a->D.cfg.hidden2 / 4);
    if (threads < tm) threads = tm;
    if (big) perform_kernel<128><<<grid, threads, perform_smem_bytes(&a->D.cfg, 128), (cudaStream_t)stream>>>(P);
    else perform_kernel<64><<<grid, threads, perform_smem_bytes(&a->D.cfg, 64), (cudaStream_t)stream>>>(P);
    ACK(cudaGetLastError());
    a->launches++;
    return REXSIM_OK;
}
int rexagent_experience(RexAgent* a, const float* observ, const float* reward, int32_t n, void* stream) {
    if (!a || !observ || !reward) return afail(REXSIM_ERR_INVALID, "null argument");
    if (n <= 0) return afail(REXSIM_ERR_INVALID, "agent: n must be positive");
    ExperienceArgs P; P.D = a->D; P.observ = observ; P.reward = reward; P.n = n; P.sums_out = nullptr;
    int blocks = (n + 255) / 256;
    if (blocks > RED_BLOCKS) blocks = RED_BLOCKS;
    experience_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P);
    ACK(cudaGetLastError());
    a->launches++;
    return REXSIM_OK;
}
int rexagent_experience_partial(RexAgent* a, const float* observ, const float* reward, int32_t n, float* sums, void* stream) {
    if (!a || !observ || !reward || !sums) return afail(REXSIM_ERR_INVALID, "null argument");
    if (n <= 0) return afail(REXSIM_ERR_INVALID, "agent: n must be positive");
    ExperienceArgs P; P.D = a->D; P.observ = observ; P.reward = reward; P.n = n; P.sums_out = sums;
    int blocks = (n + 255) / 256;
    if (blocks > RED_BLOCKS) blocks = RED_BLOCKS;
    experience_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P);
    ACK(cudaGetLastError());
    a->launches++;
    return REXSIM_OK;
}
int rexagent_experience_finalize(RexAgent* a, const float* sums, int32_t n_total, const float* observ, const float* reward, void* stream) {
    if (!a || !sums || !observ || !reward) return afail(REXSIM_ERR_INVALID, "null argument");
    if (n_total <= 0) return afail(REXSIM_ERR_INVALID, "agent: n_total must be positive");
    experience_finalize_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(a->D, sums, n_total, observ, reward);
    ACK(cudaGetLastError());
    a->launches++;
    return REXSIM_OK;
}
int rexagent_transform_reward(RexAgent* a, const float* reward, int32_t n, float* out, void* stream) {
    if (!a || !reward || !out) return afail(REXSIM_ERR_INVALID, "null argument");
    if (n <= 0) return REXSIM_OK;
    int blocks = (n + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    transform_reward_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a->D, reward, n, out);
    ACK(cudaGetLastError());
    a->launches++;
    return REXSIM_OK;
}
int rexagent_discounted_return(const float* reward, const int32_t* length, int32_t episodes, int32_t L, int64_t se, int64_t st,
                               float discount, float* out, void* stream) {
    if (!reward || !length || !out) return afail(REXSIM_ERR_INVALID, "null argument");
    if (episodes <= 0 || L <= 0) return afail(REXSIM_ERR_INVALID, "agent: empty scan");
    row_scan_kernel<false><<<(episodes + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reward, nullptr, length, episodes, L, se, st, discount, out);
    ACK(cudaGetLastError());
    return REXSIM_OK;
}
int rexagent_lambda_advantage(const float* reward, const float* value, const int32_t* length, int32_t episodes, int32_t L,
                              int64_t se, int64_t st, float discount, float* out, void* stream) {
    if (!reward || !value || !length || !out) return afail(REXSIM_ERR_INVALID, "null argument");
    if (episodes <= 0 || L <= 0) return afail(REXSIM_ERR_INVALID, "agent: empty scan");
    row_scan_kernel<true><<<(episodes + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reward, value, length, episodes, L, se, st, discount, out);
    ACK(cudaGetLastError());
    return REXSIM_OK;
}
int rexagent_gae_segments(const float* reward, const float* value, const uint8_t* done, int32_t T, int32_t n, float discount,
                          float lambda, float* out_return, float* out_advantage, void* stream) {
    if (!reward || !value || !done) return afail(REXSIM_ERR_INVALID, "null argument");
    if (T <= 0 || n <= 0) return afail(REXSIM_ERR_INVALID, "agent: empty scan");
    gae_segments_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reward, value, done, T, n, discount, lambda, out_return, out_advantage);
    ACK(cudaGetLastError());
    return REXSIM_OK;
}
int64_t rexagent_launch_count(const RexAgent* a) { return a ? a->launches : 0; }

}  // extern "C"
