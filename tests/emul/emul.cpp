// Serial CPU build of the chain engine (fthmc_b200/csrc/chain_engine.cuh) -- TEST INFRASTRUCTURE ONLY.
// The engine's phases are correct for any thread count; here they run with one "thread" so the
// indexing, the weight packing and the hand-derived adjoint can be checked against the oracle
// without a GPU.  Nothing in the product package loads this library.
#include <barrier>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../fthmc_b200/csrc/chain_programs.cuh"
#include "../../fthmc_b200/csrc/weight_pack.h"

static int g_ntraj = 1;
extern "C" void emul_set_ntraj(int n) { g_ntraj = n < 1 ? 1 : n; }   // trajectories per emul_run call (run-loop mode)

namespace {
struct SerialExec {
    static constexpr bool kCluster = false;
    int rank() const { return 0; }
    int nranks() const { return 1; }
    double* peer(double* p, int) const { return p; }
    double* arena;
    double* smem() const { return arena; }
    void bar_init(int) const {}
    void bulk_load(int, double* dst, const double* src, int n) const { memcpy(dst, src, sizeof(double) * n); }
    void bar_wait(int, int) const {}
    void proxy_fence() const {}
    int tid() const { return 0; }
    int nt() const { return 1; }
    bool fine(int) const { return getenv("FT_EMUL_FINE") != nullptr; }     // exercise the finer task split of wide blocks
    // warp-level tile D(8x8) += A(8x4) B(4x8) in the PTX m8n8k4 fragment layout; this "thread" carries all 32 lanes
    static constexpr int kLanes = 32;
    bool use_mma() const { return getenv("FT_EMUL_MMA") != nullptr; }
    int warp() const { return 0; }
    int nwarps() const { return 1; }
    int lane0() const { return 0; }
    void mma884(double (&d0)[32], double (&d1)[32], const double (&a)[32], const double (&b)[32]) const {
        for (int i = 0; i < 8; ++i)
            for (int n = 0; n < 8; ++n) {
                double& d = (n & 1) ? d1[4 * i + (n >> 1)] : d0[4 * i + (n >> 1)];
                for (int j = 0; j < 4; ++j) d = fma(a[4 * i + j], b[4 * n + j], d);
            }
    }
    void sync() const {}
    void lsync() const {}
    double sum(double v) const { return v; }
    double maxv(double v) const { return v; }
    bool all(bool pred) const { return pred; }
};
}

extern "C" int emul_run(int mode, int B, int L0, int L1, int nlayers, const double* raw, const int* mu, const int* off,
                        int act, int conv, double tol, int max_iter, double beta, double dt, int nstep,
                        const double* field_in, const double* p_in, const double* u_in, double* field_out, double* p_out,
                        double* s_out, double* layer_logJ, int* iters, double* expmdH, int* acc, double* plaq, double* topo,
                        double* h0, double* h1, unsigned long long seed, unsigned long long traj) {
    using namespace fthmc;
    if (L0 % 4 || L1 % 4) return -1;
    std::vector<double> pack((size_t)nlayers * PACK_DOUBLES);
    for (int l = 0; l < nlayers; ++l) pack_layer(raw + (size_t)l * RAW_DOUBLES, mu[l], pack.data() + (size_t)l * PACK_DOUBLES);
    std::vector<double> smem(engine_smem_doubles(L0, L1, nlayers > 0) + 8), ws(engine_ws_doubles(L0, L1, nlayers) + 8);
    ChainArgs a{};
    a.mode = mode; a.B = B;
    a.pr.L0 = L0; a.pr.L1 = L1; a.pr.nlayers = nlayers; a.pr.act = act; a.pr.conv = conv;
    a.pr.inv_tol = tol; a.pr.inv_max_iter = max_iter; a.pr.wpack = pack.data(); a.pr.lmu = mu; a.pr.loff = off;
    a.beta = beta; a.dt = dt; a.nstep = nstep;
    a.field_in = field_in; a.p_in = p_in; a.u_in = u_in; a.field_out = field_out; a.p_out = p_out;
    a.s_out = s_out; a.layer_logJ = layer_logJ; a.iters = iters;
    a.expmdH = expmdH; a.acc = acc; a.plaq = plaq; a.topo = topo; a.h0 = h0; a.h1 = h1;
    a.seed = seed; a.traj = traj; a.chain0 = 0; a.ntraj = g_ntraj;
    SerialExec ex{ smem.data() };
    Engine<SerialExec> en(ex, a.pr, ws.data());
    if (nlayers > 0) en.load_geom_table();
    for (int b = 0; b < B; ++b) run_chain(en, a, b);
    return 0;
}


// Training gradient (MODE_FT_GRAD): action (B) and d/d(raw weights) of sum_b ft_action(x_b), (nlayers, 955).
// Needs the tensor-core code path (FT_EMUL_MMA) and L0, L1 multiples of 8.
static const double* g_vjp_seed = nullptr;       // set by emul_set_vjp: the next emul_grad call runs the vector-Jacobian mode
static const double* g_vjp_wlj = nullptr;
extern "C" void emul_set_vjp(const double* seed, const double* wlj) { g_vjp_seed = seed; g_vjp_wlj = wlj; }

extern "C" int emul_grad(int B, int L0, int L1, int nlayers, const double* raw, const int* mu, const int* off,
                         int act, int conv, double beta, const double* field_in, double* action_out, double* grad_raw, double* force_out) {
    using namespace fthmc;
    if (L0 % 8 || L1 % 8 || nlayers <= 0) return -1;
    std::vector<double> pack((size_t)nlayers * PACK_DOUBLES);
    for (int l = 0; l < nlayers; ++l) pack_layer(raw + (size_t)l * RAW_DOUBLES, mu[l], pack.data() + (size_t)l * PACK_DOUBLES);
    std::vector<double> smem(engine_smem_doubles(L0, L1, true) + 8), ws(engine_ws_doubles(L0, L1, nlayers, 1, true) + 8);
    std::vector<double> gbuf((size_t)nlayers * GRAD_DOUBLES, 0.0);
    ChainArgs a{};
    a.mode = MODE_FT_GRAD; a.B = B; a.ntraj = 1;
    a.pr.L0 = L0; a.pr.L1 = L1; a.pr.nlayers = nlayers; a.pr.act = act; a.pr.conv = conv;
    a.pr.inv_tol = 1e-6; a.pr.inv_max_iter = 1000; a.pr.wpack = pack.data(); a.pr.lmu = mu; a.pr.loff = off; a.pr.train = 1;
    a.beta = beta; a.field_in = field_in; a.field_out = force_out; a.s_out = action_out;
    a.vjp_seed = g_vjp_seed; a.vjp_wlj = g_vjp_wlj;
    g_vjp_seed = nullptr; g_vjp_wlj = nullptr;
    SerialExec ex{ smem.data() };
    if (!ex.use_mma()) return -3;
    Engine<SerialExec> en(ex, a.pr, ws.data());
    en.gW = gbuf.data();
    en.load_geom_table();
    for (int b = 0; b < B; ++b) run_chain(en, a, b);
    for (int l = 0; l < nlayers; ++l) unpack_grad_layer(gbuf.data() + (size_t)l * GRAD_DOUBLES, mu[l], grad_raw + (size_t)l * RAW_DOUBLES);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Cluster mode: nr host threads stand in for the nr CTAs of a thread-block cluster (one "thread" per
// CTA), std::barrier for the cluster barrier, plain pointers into the other ranks' arenas for
// distributed shared memory.  Exercises the decomposition (row-block X/GR, column-block planes,
// halo pushes, cluster-wide reductions) of the kCluster code path without a GPU.
// ------------------------------------------------------------------------------------------------
namespace {
struct ClusterShared {
    int nr;
    std::vector<double*> arenas;
    std::barrier<> bar;
    std::vector<double> slots;
    ClusterShared(int n) : nr(n), arenas(n), bar(n), slots(n) {}
};
struct ThreadExec {
    static constexpr bool kCluster = true;
    ClusterShared* sh;
    int rk;
    int rank() const { return rk; }
    int nranks() const { return sh->nr; }
    double* smem() const { return sh->arenas[rk]; }
    double* peer(double* p, int r) const { return sh->arenas[r] + (p - sh->arenas[rk]); }
    void bar_init(int) const {}
    void bulk_load(int, double* dst, const double* src, int n) const { memcpy(dst, src, sizeof(double) * n); }
    void bar_wait(int, int) const {}
    void proxy_fence() const {}
    int tid() const { return 0; }
    int nt() const { return 1; }
    bool fine(int) const { return getenv("FT_EMUL_FINE") != nullptr; }
    // warp-level tile D(8x8) += A(8x4) B(4x8) in the PTX m8n8k4 fragment layout; this "thread" carries all 32 lanes
    static constexpr int kLanes = 32;
    bool use_mma() const { return getenv("FT_EMUL_MMA") != nullptr; }
    int warp() const { return 0; }
    int nwarps() const { return 1; }
    int lane0() const { return 0; }
    void mma884(double (&d0)[32], double (&d1)[32], const double (&a)[32], const double (&b)[32]) const {
        for (int i = 0; i < 8; ++i)
            for (int n = 0; n < 8; ++n) {
                double& d = (n & 1) ? d1[4 * i + (n >> 1)] : d0[4 * i + (n >> 1)];
                for (int j = 0; j < 4; ++j) d = fma(a[4 * i + j], b[4 * n + j], d);
            }
    }
    void sync() const { sh->bar.arrive_and_wait(); }
    void lsync() const {}          // one "thread" per CTA: a CTA-local barrier is a no-op
    double sum(double v) const {
        sh->slots[rk] = v; sh->bar.arrive_and_wait();
        double t = 0.0; for (int i = 0; i < sh->nr; ++i) t += sh->slots[i];
        sh->bar.arrive_and_wait();
        return t;
    }
    bool all(bool pred) const { return maxv(pred ? 0.0 : 1.0) == 0.0; }
    double maxv(double v) const {
        sh->slots[rk] = v; sh->bar.arrive_and_wait();
        double t = sh->slots[0]; for (int i = 1; i < sh->nr; ++i) t = t > sh->slots[i] ? t : sh->slots[i];
        sh->bar.arrive_and_wait();
        return t;
    }
};
}

extern "C" int emul_run_cluster(int nr, int mode, int B, int L0, int L1, int nlayers, const double* raw, const int* mu, const int* off,
                                int act, int conv, double tol, int max_iter, double beta, double dt, int nstep,
                                const double* field_in, const double* p_in, const double* u_in, double* field_out, double* p_out,
                                double* s_out, double* layer_logJ, int* iters, double* expmdH, int* acc, double* plaq, double* topo,
                                double* h0, double* h1, unsigned long long seed, unsigned long long traj) {
    using namespace fthmc;
    if (L0 % 4 || L1 % 4 || nr < 1 || (L0 / 4) % nr || (L1 / 4) % nr) return -1;
    if (nlayers > 0 && (size_t)L0 * L1 / nr > (size_t)OFF_W3T) return -2;
    std::vector<double> pack((size_t)nlayers * PACK_DOUBLES);
    for (int l = 0; l < nlayers; ++l) pack_layer(raw + (size_t)l * RAW_DOUBLES, mu[l], pack.data() + (size_t)l * PACK_DOUBLES);
    const size_t asz = engine_smem_doubles(L0, L1, nlayers > 0, nr) + 8;
    std::vector<std::vector<double>> arenas(nr, std::vector<double>(asz));
    std::vector<double> ws(engine_ws_doubles(L0, L1, nlayers, nr) + 8);
    ChainArgs a{};
    a.mode = mode; a.B = B;
    a.pr.L0 = L0; a.pr.L1 = L1; a.pr.nlayers = nlayers; a.pr.act = act; a.pr.conv = conv;
    a.pr.inv_tol = tol; a.pr.inv_max_iter = max_iter; a.pr.wpack = pack.data(); a.pr.lmu = mu; a.pr.loff = off;
    a.beta = beta; a.dt = dt; a.nstep = nstep;
    a.field_in = field_in; a.p_in = p_in; a.u_in = u_in; a.field_out = field_out; a.p_out = p_out;
    a.s_out = s_out; a.layer_logJ = layer_logJ; a.iters = iters;
    a.expmdH = expmdH; a.acc = acc; a.plaq = plaq; a.topo = topo; a.h0 = h0; a.h1 = h1;
    a.seed = seed; a.traj = traj; a.chain0 = 0; a.ntraj = g_ntraj;
    ClusterShared sh(nr);
    for (int r = 0; r < nr; ++r) sh.arenas[r] = arenas[r].data();
    std::vector<std::thread> th;
    for (int r = 0; r < nr; ++r)
        th.emplace_back([&, r]() {
            ThreadExec ex{ &sh, r };
            Engine<ThreadExec> en(ex, a.pr, ws.data());
            if (nlayers > 0) en.load_geom_table();
            for (int b = 0; b < B; ++b) run_chain(en, a, b);
        });
    for (auto& t : th) t.join();
    return 0;
}
