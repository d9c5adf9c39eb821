// A/B microbenchmark: lane-per-chain leapfrog of the radon-919 shape against the warp-per-chain
// kernel of the product path (csrc/pm4_warp.cuh).  NOT on the product path: a measuring tool, built and run by hand
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_bin/lane_ab tools/lane_ab.cu
//   tools/_bin/lane_ab > gpurun_out/lane_ab.jsonl
//
// Layout measured here: thread = chain, CTA = 32 chains x W warps.  The chain state (theta, p, grad: 3 x D floats per
// chain) sits chain-minor in shared memory ([d][lane]: bank = lane, conflict-free), the observation records (x, y)
// are read as broadcasts (shared memory or the constant bank).  The G groups are dealt to the W warps of the CTA in
// contiguous ranges of equal cost: a group's likelihood sums need no cross-lane reduction at all; only the five
// scalar adjoints (mu_a, log sigma_a, mu_b, log sigma_b, log eps) cross the warps (W partial sums per chain through
// shared memory, summed by warp 0).  Two CTA barriers per leapfrog.
//
// The model is the centred radon model of pymc4_b200/benchmodels.py::hierarchical_model in unconstrained
// coordinates (theta order of the oracle: mu_a, log s_a, mu_b, log s_b, a[G], b[G], log eps), gradient in closed form
// (SURVEY App. B).  One unit = one leapfrog (gradient + kick + drift) of one chain, the unit of BASELINE.json's metric;
// there is no tree logic here, so the number is an UPPER bound for a NUTS kernel in this layout (a warp walks the
// deepest of its 32 trees: see DESIGN.md section 3.0 for the masked fraction).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <algorithm>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int G = 85, N = 919, D = 2 * G + 5;
constexpr int MAXW = 32;

__constant__ float2 c_obs[N];          // (x, y) per observation, sorted by group
__constant__ int c_gstart[G + 1];      // first observation of every group
__constant__ int c_wstart[MAXW + 1];   // first group of every warp (filled per W before a launch)

template <int W, bool OBS_CONST, bool MASKED>
__global__ void __launch_bounds__(32 * W) lane_leapfrog(const float* __restrict__ th0, const float* __restrict__ p0,
                                                       float* __restrict__ th_out, float* __restrict__ p_out,
                                                       int C, int L, float eps, const int* __restrict__ Lc) {
    extern __shared__ float sm[];
    float* th = sm;
    float* p = th + D * 32;
    float* gr = p + D * 32;
    float* red = gr + D * 32;                        // [W][5][32]
    float2* obs = reinterpret_cast<float2*>(red + W * 5 * 32);   // [N] (only when !OBS_CONST)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int chain = blockIdx.x * 32 + lane;
    for (int d = w; d < D; d += W) {
        th[d * 32 + lane] = th0[(size_t)d * C + chain];
        p[d * 32 + lane] = p0[(size_t)d * C + chain];
    }
    if (!OBS_CONST)
        for (int i = threadIdx.x; i < N; i += 32 * W) obs[i] = c_obs[i];
    __syncthreads();
    const int gs = c_wstart[w], ge = c_wstart[w + 1];
    // `Lc` (masked mode): every chain walks its own number of leapfrogs, the CTA walks the longest of its 32 chains (what
    // a NUTS kernel in this layout does with trees of different sizes); every warp of the CTA sees the same 32 chains
    const int myL = MASKED ? Lc[chain] : L;
    if (MASKED) L = __reduce_max_sync(0xffffffffu, myL);
    for (int step = 0; step < L; ++step) {
        const bool act = !MASKED || step < myL;
        const float mua = th[0 * 32 + lane], lsa = th[1 * 32 + lane], mub = th[2 * 32 + lane], lsb = th[3 * 32 + lane];
        const float le = th[(D - 1) * 32 + lane];
        const float isa2 = __expf(-2.f * lsa), isb2 = __expf(-2.f * lsb), ie2 = __expf(-2.f * le);
        float acc_ma = 0.f, acc_za = 0.f, acc_mb = 0.f, acc_zb = 0.f, acc_s2 = 0.f;
        for (int g = gs; g < ge; ++g) {
            const float a = th[(4 + g) * 32 + lane], b = th[(4 + G + g) * 32 + lane];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            const int i0 = c_gstart[g], i1 = c_gstart[g + 1];
#pragma unroll 4
            for (int i = i0; i < i1; ++i) {
                const float2 o = OBS_CONST ? c_obs[i] : obs[i];
                const float d = fmaf(-o.x, b, o.y) - a;
                s0 += d;
                s1 = fmaf(d, o.x, s1);
                s2 = fmaf(d, d, s2);
            }
            const float da = a - mua, db = b - mub;
            const float ga = fmaf(s0, ie2, -da * isa2), gb = fmaf(s1, ie2, -db * isb2);
            acc_ma += da; acc_za = fmaf(da, da, acc_za);
            acc_mb += db; acc_zb = fmaf(db, db, acc_zb);
            acc_s2 += s2;
            // kick + drift of the owned coordinates right away (nobody else reads them)
            const float pa = fmaf(eps, ga, p[(4 + g) * 32 + lane]), pb = fmaf(eps, gb, p[(4 + G + g) * 32 + lane]);
            if (act) {
                p[(4 + g) * 32 + lane] = pa; p[(4 + G + g) * 32 + lane] = pb;
                gr[(4 + g) * 32 + lane] = ga; gr[(4 + G + g) * 32 + lane] = gb;
                th[(4 + g) * 32 + lane] = fmaf(eps, pa, a); th[(4 + G + g) * 32 + lane] = fmaf(eps, pb, b);
            }
        }
        red[(w * 5 + 0) * 32 + lane] = acc_ma * isa2;
        red[(w * 5 + 1) * 32 + lane] = acc_za * isa2;
        red[(w * 5 + 2) * 32 + lane] = acc_mb * isb2;
        red[(w * 5 + 3) * 32 + lane] = acc_zb * isb2;
        red[(w * 5 + 4) * 32 + lane] = acc_s2 * ie2;
        __syncthreads();
        if (w == 0) {
            float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int v = 0; v < W; ++v)
#pragma unroll
                for (int k = 0; k < 5; ++k) t[k] += red[(v * 5 + k) * 32 + lane];
            const float sa2 = __expf(2.f * lsa), sb2 = __expf(2.f * lsb), e2 = __expf(2.f * le);
            float gsc[5];
            gsc[0] = t[0] - mua;
            gsc[1] = t[1] - (float)G + 1.f - 2.f * sa2 / (1.f + sa2);
            gsc[2] = t[2] - mub;
            gsc[3] = t[3] - (float)G + 1.f - 2.f * sb2 / (1.f + sb2);
            gsc[4] = t[4] - (float)N + 1.f - 2.f * e2 / (1.f + e2);
            const int dd[5] = {0, 1, 2, 3, D - 1};
            const float cur[5] = {mua, lsa, mub, lsb, le};
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float pk = fmaf(eps, gsc[k], p[dd[k] * 32 + lane]);
                if (act) {
                    p[dd[k] * 32 + lane] = pk;
                    gr[dd[k] * 32 + lane] = gsc[k];
                    th[dd[k] * 32 + lane] = fmaf(eps, pk, cur[k]);
                }
            }
        }
        __syncthreads();
    }
    for (int d = w; d < D; d += W) {
        th_out[(size_t)d * C + chain] = th[d * 32 + lane];
        p_out[(size_t)d * C + chain] = p[d * 32 + lane];
    }
}

// the same recurrence in double on the host, one chain
static void host_chain(const std::vector<float2>& obs, const std::vector<int>& gstart, std::vector<double>& th,
                       std::vector<double>& p, int L, double eps) {
    for (int step = 0; step < L; ++step) {
        const double mua = th[0], lsa = th[1], mub = th[2], lsb = th[3], le = th[D - 1];
        const double isa2 = exp(-2 * lsa), isb2 = exp(-2 * lsb), ie2 = exp(-2 * le);
        double t[5] = {0, 0, 0, 0, 0};
        for (int g = 0; g < G; ++g) {
            const double a = th[4 + g], b = th[4 + G + g];
            double s0 = 0, s1 = 0, s2 = 0;
            for (int i = gstart[g]; i < gstart[g + 1]; ++i) {
                const double d = obs[i].y - a - b * obs[i].x;
                s0 += d; s1 += d * obs[i].x; s2 += d * d;
            }
            const double da = a - mua, db = b - mub;
            const double ga = s0 * ie2 - da * isa2, gb = s1 * ie2 - db * isb2;
            t[0] += da * isa2; t[1] += da * da * isa2; t[2] += db * isb2; t[3] += db * db * isb2; t[4] += s2 * ie2;
            p[4 + g] += eps * ga; p[4 + G + g] += eps * gb;
            th[4 + g] = a + eps * p[4 + g]; th[4 + G + g] = b + eps * p[4 + G + g];
        }
        const double sa2 = exp(2 * lsa), sb2 = exp(2 * lsb), e2 = exp(2 * le);
        const double gsc[5] = {t[0] - mua, t[1] - G + 1 - 2 * sa2 / (1 + sa2), t[2] - mub,
                               t[3] - G + 1 - 2 * sb2 / (1 + sb2), t[4] - N + 1 - 2 * e2 / (1 + e2)};
        const int dd[5] = {0, 1, 2, 3, D - 1};
        for (int k = 0; k < 5; ++k) { p[dd[k]] += eps * gsc[k]; th[dd[k]] += eps * p[dd[k]]; }
    }
}

template <int W, bool OC>
static void run_one(int C, int L, float eps, const float* d_th0, const float* d_p0, float* d_th, float* d_p,
                    const std::vector<int>& gstart, const std::vector<float>& h_th0, const std::vector<float>& h_p0,
                    const std::vector<float2>& obs) {
    // contiguous group ranges of equal cost (7 per observation + 35 per group)
    std::vector<int> ws(MAXW + 1, G);
    {
        double total = 7.0 * N + 35.0 * G, acc = 0; int w = 0; ws[0] = 0;
        for (int g = 0; g < G; ++g) {
            acc += 7.0 * (gstart[g + 1] - gstart[g]) + 35.0;
            while (w + 1 < W && acc >= total * (w + 1) / W) ws[++w] = g + 1;
        }
        for (int v = w + 1; v <= MAXW; ++v) ws[v] = G;
    }
    CK(cudaMemcpyToSymbol(c_wstart, ws.data(), sizeof(int) * (MAXW + 1)));
    const size_t smem = sizeof(float) * (3 * D * 32 + W * 5 * 32) + (OC ? 0 : sizeof(float2) * N);
    auto kern = lane_leapfrog<W, OC, false>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, ms[4];
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<C / 32, 32 * W, smem>>>(d_th0, d_p0, d_th, d_p, C, L, eps, nullptr);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms[rep], e0, e1));
        if (rep > 0) best = std::min(best, ms[rep]);
    }
    // parity of three chains against the double recurrence
    std::vector<float> h_th((size_t)D * C);
    CK(cudaMemcpy(h_th.data(), d_th, sizeof(float) * D * C, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    const int probe[3] = {0, C / 2 + 7, C - 1};
    for (int c : probe) {
        std::vector<double> th(D), p(D);
        for (int d = 0; d < D; ++d) { th[d] = h_th0[(size_t)d * C + c]; p[d] = h_p0[(size_t)d * C + c]; }
        host_chain(obs, gstart, th, p, L, eps);
        for (int d = 0; d < D; ++d) maxerr = std::max(maxerr, fabs(th[d] - (double)h_th[(size_t)d * C + c]));
    }
    printf("{\"layout\": \"lane-per-chain\", \"chains\": %d, \"warps_per_cta\": %d, \"obs\": \"%s\", \"leapfrogs\": %d, "
           "\"ms\": %.4f, \"units_per_s\": %.4e, \"us_per_leapfrog_of_a_cta\": %.3f, \"ctas_per_sm\": %d, "
           "\"smem_bytes\": %zu, \"max_abs_err_vs_f64\": %.3e}\n",
           C, W, OC ? "constant bank" : "shared broadcast", L, best, (double)C * L / (best * 1e-3),
           best * 1e3 / L / std::max(1.0, ceil((double)(C / 32) / (148.0 * occ))), occ, smem, maxerr);
    fflush(stdout);
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

// Masked mode: tree sizes drawn from the distribution measured on the headline run (tools/tree_stats.py: 77 % of the
// trees 31 leapfrogs, 15 % 63, 0.12 % 127, 0.011 % 255, the rest 15), one launch = one lock-step transition of all
// chains in this layout (its duration is set by the CTA with the deepest tree); no tree bookkeeping is executed.
template <int W>
static void run_masked(int C, float eps, const float* d_th0, const float* d_p0, float* d_th, float* d_p,
                       const std::vector<int>& gstart, const std::vector<float>& h_th0, const std::vector<float>& h_p0,
                       const std::vector<float2>& obs, std::mt19937_64& rng) {
    std::vector<int> ws(MAXW + 1, G);
    {
        double total = 7.0 * N + 35.0 * G, acc = 0; int w = 0; ws[0] = 0;
        for (int g = 0; g < G; ++g) {
            acc += 7.0 * (gstart[g + 1] - gstart[g]) + 35.0;
            while (w + 1 < W && acc >= total * (w + 1) / W) ws[++w] = g + 1;
        }
        for (int v = w + 1; v <= MAXW; ++v) ws[v] = G;
    }
    CK(cudaMemcpyToSymbol(c_wstart, ws.data(), sizeof(int) * (MAXW + 1)));
    const size_t smem = sizeof(float) * (3 * D * 32 + W * 5 * 32) + sizeof(float2) * N;
    auto kern = lane_leapfrog<W, false, true>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int* d_L; CK(cudaMalloc(&d_L, sizeof(int) * C));
    std::uniform_real_distribution<double> ud(0.0, 1.0);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int R = 24;
    double sum_ms = 0, max_ms = 0, useful = 0, walked = 0, maxerr = 0; int deepest = 0;
    std::vector<int> hL(C);
    std::vector<float> h_th((size_t)D * C);
    for (int rep = 0; rep < R + 1; ++rep) {
        for (int c = 0; c < C; ++c) {
            const double u = ud(rng);
            hL[c] = u < 0.77 ? 31 : u < 0.92 ? 63 : u < 0.9212 ? 127 : u < 0.92131 ? 255 : 15;
        }
        CK(cudaMemcpy(d_L, hL.data(), sizeof(int) * C, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0));
        kern<<<C / 32, 32 * W, smem>>>(d_th0, d_p0, d_th, d_p, C, 0, eps, d_L);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 0) continue;   // warm-up
        sum_ms += ms; max_ms = std::max(max_ms, (double)ms);
        for (int b = 0; b < C / 32; ++b) {
            int mx = 0;
            for (int l = 0; l < 32; ++l) { useful += hL[b * 32 + l]; mx = std::max(mx, hL[b * 32 + l]); }
            walked += 32.0 * mx; deepest = std::max(deepest, mx);
        }
        if (rep == 1) {
            CK(cudaMemcpy(h_th.data(), d_th, sizeof(float) * D * C, cudaMemcpyDeviceToHost));
            const int probe[3] = {0, C / 2 + 7, C - 1};
            for (int c : probe) {
                std::vector<double> th(D), p(D);
                for (int d = 0; d < D; ++d) { th[d] = h_th0[(size_t)d * C + c]; p[d] = h_p0[(size_t)d * C + c]; }
                host_chain(obs, gstart, th, p, hL[c], eps);
                for (int d = 0; d < D; ++d) maxerr = std::max(maxerr, fabs(th[d] - (double)h_th[(size_t)d * C + c]));
            }
        }
    }
    printf("{\"layout\": \"lane-per-chain, tree sizes of the headline run (masked lanes)\", \"chains\": %d, \"warps_per_cta\": %d, "
           "\"launches\": %d, \"mean_ms_per_transition\": %.4f, \"max_ms_per_transition\": %.4f, \"useful_units_per_s\": %.4e, "
           "\"useful_lane_fraction\": %.4f, \"deepest_tree\": %d, \"max_abs_err_vs_f64\": %.3e}\n",
           C, W, R, sum_ms / R, max_ms, useful / (sum_ms * 1e-3), useful / walked, deepest, maxerr);
    fflush(stdout);
    CK(cudaFree(d_L)); CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

int main(int argc, char** argv) {
    // `lane_ab one C W`: a single configuration (shared-memory broadcast), for ncu
    const bool one = argc >= 4 && std::string(argv[1]) == "one";
    const int oneC = one ? atoi(argv[2]) : 0, oneW = one ? atoi(argv[3]) : 0;
    std::mt19937_64 rng(20261019);
    // radon-shaped group sizes: every group at least one observation, log-normal weights
    std::vector<double> wgt(G);
    std::lognormal_distribution<double> ln(0.0, 1.0);
    for (auto& v : wgt) v = ln(rng);
    std::discrete_distribution<int> pick(wgt.begin(), wgt.end());
    std::vector<int> sizes(G, 1);
    for (int i = 0; i < N - G; ++i) sizes[pick(rng)]++;
    std::vector<int> gstart(G + 1, 0);
    for (int g = 0; g < G; ++g) gstart[g + 1] = gstart[g] + sizes[g];
    std::normal_distribution<double> nd(0.0, 1.0);
    std::uniform_real_distribution<double> ud(0.0, 1.0);
    std::vector<float2> obs(N);
    for (int g = 0; g < G; ++g) {
        const double a = 1.5 + 0.33 * nd(rng), b = -0.65 + 0.35 * nd(rng);
        for (int i = gstart[g]; i < gstart[g + 1]; ++i) {
            const double x = ud(rng) < 0.17 ? 1.0 : 0.0;
            obs[i] = make_float2((float)x, (float)(a + b * x + 0.72 * nd(rng)));
        }
    }
    CK(cudaMemcpyToSymbol(c_obs, obs.data(), sizeof(float2) * N));
    CK(cudaMemcpyToSymbol(c_gstart, gstart.data(), sizeof(int) * (G + 1)));
    const int L = 64;
    const float eps = 0.004f;
    const int Cs[3] = {4096, 16384, 65536};
    for (int C : Cs) {
        if (one && C != oneC) continue;
        std::vector<float> h_th0((size_t)D * C), h_p0((size_t)D * C);
        for (int c = 0; c < C; ++c)
            for (int d = 0; d < D; ++d) {
                double v = 0.1 * nd(rng);
                if (d == 0 || (d >= 4 && d < 4 + G)) v += 1.5;
                if (d == 2 || (d >= 4 + G && d < 4 + 2 * G)) v -= 0.65;
                if (d == 1 || d == 3) v += log(0.35);
                if (d == D - 1) v += log(0.72);
                h_th0[(size_t)d * C + c] = (float)v;
                h_p0[(size_t)d * C + c] = (float)nd(rng);
            }
        float *d_th0, *d_p0, *d_th, *d_p;
        const size_t bytes = sizeof(float) * D * C;
        CK(cudaMalloc(&d_th0, bytes)); CK(cudaMalloc(&d_p0, bytes)); CK(cudaMalloc(&d_th, bytes)); CK(cudaMalloc(&d_p, bytes));
        CK(cudaMemcpy(d_th0, h_th0.data(), bytes, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_p0, h_p0.data(), bytes, cudaMemcpyHostToDevice));
#define RUN(W, OC) run_one<W, OC>(C, L, eps, d_th0, d_p0, d_th, d_p, gstart, h_th0, h_p0, obs)
        if (one) {
            if (oneW == 4) RUN(4, false); else if (oneW == 8) RUN(8, false); else if (oneW == 32) RUN(32, false); else RUN(16, false);
        } else {
        RUN(1, false); RUN(2, false); RUN(4, false); RUN(8, false); RUN(16, false); RUN(32, false);
        RUN(1, true); RUN(4, true); RUN(8, true); RUN(16, true);
        run_masked<8>(C, eps, d_th0, d_p0, d_th, d_p, gstart, h_th0, h_p0, obs, rng);
        run_masked<16>(C, eps, d_th0, d_p0, d_th, d_p, gstart, h_th0, h_p0, obs, rng);
        run_masked<32>(C, eps, d_th0, d_p0, d_th, d_p, gstart, h_th0, h_p0, obs, rng);
        }
#undef RUN
        CK(cudaFree(d_th0)); CK(cudaFree(d_p0)); CK(cudaFree(d_th)); CK(cudaFree(d_p));
    }
    return 0;
}
