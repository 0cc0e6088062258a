// chain_programs.cuh -- what one CTA does with one chain for each C-ABI entry point.
// Shared by the CUDA kernel (fthmc_capi.cu) and the serial CPU emulation used only by tests.
#pragma once
#include "chain_engine.cuh"

namespace fthmc {

enum ChainMode {
    MODE_FLOW_FWD = 0,     // ft_flow         ipynb/ft_hmc.py:220   (+ summed / per-layer logJ)
    MODE_FLOW_INV = 1,     // ft_flow_inv     ipynb/ft_hmc.py:225
    MODE_FT_ACTION = 2,    // ft_action       ipynb/ft_hmc.py:230
    MODE_FT_FORCE = 3,     // ft_force        ipynb/ft_hmc.py:240
    MODE_FT_LEAPFROG = 4,  // ft_leapfrog     ipynb/ft_hmc.py:394
    MODE_FT_HMC = 5,       // ft_hmc          ipynb/ft_hmc.py:420
    MODE_HMC = 6,          // hmc             hmc_2dU1.py:144
    MODE_LEAPFROG = 7,     // leapfrog        hmc_2dU1.py:132
    MODE_FT_GRAD = 8,      // d/dweights of sum_b ft_action(x_b): the reverse-KL training gradient, ipynb/ft_hmc.py:253-295
};

struct ChainArgs {
    int mode;
    int B;
    int b_begin, b_end;       // chains [b_begin, b_end) of the batch belong to this launch (set by the launcher)
    EngineParams pr;
    double beta, dt;
    int nstep;
    int ntraj;                // trajectory modes: trajectories per launch (>= 1); per-trajectory arrays are (ntraj, B, ...)
    const double* field_in;   // (B,2,L0,L1)
    const double* p_in;       // (B,2,L0,L1) or null; trajectory modes: (ntraj,B,2,L0,L1)
    const double* u_in;       // (B) or null; trajectory modes: (ntraj,B)
    double* field_out;        // (B,2,L0,L1): flowed field / force / new field
    double* p_out;            // (B,2,L0,L1) or null
    double* s_out;            // (B): logJ / action / dH
    double* layer_logJ;       // (B,nlayers) or null
    int* iters;               // (B,nlayers) or null (bisection iteration counts)
    double* expmdH; int* acc; double* plaq; double* topo; double* h0; double* h1;   // (B) each, trajectory modes
    uint64_t seed, traj, chain0;
    const double* vjp_seed;   // MODE_FT_GRAD, optional (B,2,L0,L1): external d/dy seeding the adjoint sweep instead of the Wilson force
    const double* vjp_wlj;    // MODE_FT_GRAD, optional (B): weight of sum logJ per chain (-1 for ft_action)
    double* gbuf;             // MODE_FT_GRAD: gradient accumulators, one slice of gbuf_stride doubles per CTA
    size_t gbuf_stride;
    double* ws;               // per-CTA workspace base
    size_t ws_stride;         // doubles per CTA
};

// `en` is constructed once per CTA (in shared memory on the device) by the caller
template <class E>
FT_HD void run_chain(Engine<E>& en, const ChainArgs& a, int b) {
    E& ex = en.ex;
    const size_t fs = (size_t)2 * en.Vg;
    const double* fin = a.field_in + (size_t)b * fs;
    double* fout = a.field_out ? a.field_out + (size_t)b * fs : nullptr;
    ex.sync();
    if (ex.tid() == 0) en.iters_out = a.iters ? a.iters + (size_t)b * a.pr.nlayers : nullptr;
    ex.sync();
    double* llj = a.layer_logJ ? a.layer_logJ + (size_t)b * a.pr.nlayers : nullptr;
    switch (a.mode) {
    case MODE_FLOW_FWD: {
        en.load_field(en.oX, fin); ex.sync();
        double lj = en.flow_forward(a.s_out != nullptr || llj != nullptr, false, llj);
        en.store_field(fout, en.oX);
        if (a.s_out && ex.tid() == 0) a.s_out[b] = lj;
        ex.sync();
    } break;
    case MODE_FLOW_INV: {
        en.load_field(en.oX, fin); ex.sync();
        double lj = en.flow_reverse(a.s_out != nullptr || llj != nullptr, llj);
        en.store_field(fout, en.oX);
        if (a.s_out && ex.tid() == 0) a.s_out[b] = lj;
        ex.sync();
    } break;
    case MODE_FT_ACTION: {
        en.load_field(en.oX, fin); ex.sync();
        double s = en.ft_action(a.beta);
        if (ex.tid() == 0) a.s_out[b] = s;
        if (fout) en.store_field(fout, en.oX);
        ex.sync();
    } break;
    case MODE_FT_FORCE: {
        en.load_field(en.oX, fin); ex.sync();
        en.ft_force(a.beta);
        en.store_field(fout, en.oGR);
        ex.sync();
    } break;
    case MODE_FT_GRAD: {
        if (ex.tid() == 0) {
            en.vjp_seed = a.vjp_seed ? a.vjp_seed + (size_t)b * fs : nullptr;
            en.mw = a.vjp_wlj ? -a.vjp_wlj[b] : 1.0;
        }
        en.load_field(en.oX, fin); ex.sync();
        double s = en.template ft_force<true>(a.beta, true); // weight gradients accumulate into en.gW along the adjoint sweep
        if (a.s_out && ex.tid() == 0) a.s_out[b] = s;
        if (fout) en.store_field(fout, en.oGR);              // optional: the force on the input field
        ex.sync();
    } break;
    case MODE_FT_LEAPFROG:
    case MODE_LEAPFROG: {
        en.load_field(en.oX, fin);
        en.for_links([&](int, int gi) { en.wsP[gi] = a.p_in[(size_t)b * fs + gi]; });
        ex.sync();
        if (a.mode == MODE_FT_LEAPFROG)
            leapfrog_resident(en, a.dt, a.nstep, en.wsP, [&]() { en.ft_force(a.beta); });
        else
            if constexpr (E::kCluster) leapfrog_resident(en, a.dt, a.nstep, en.wsP, [&]() { en.wilson_force(a.beta, 1); });
            else leapfrog_plain_fused(en, a.beta, a.dt, a.nstep, en.wsP);
        en.store_field(fout, en.oX);
        en.for_links([&](int, int gi) { a.p_out[(size_t)b * fs + gi] = en.wsP[gi]; });
        ex.sync();
    } break;
    case MODE_FT_HMC:
    case MODE_HMC: {
        // run loops (ipynb/ft_hmc.py:180, 437; hmc_2dU1.py:697): ntraj trajectories of this chain in one launch, the
        // field resident in shared memory throughout; row t of every per-trajectory array belongs to trajectory t
        const int nt = a.ntraj < 1 ? 1 : a.ntraj;
        for (int t = 0; t < nt; ++t) {
            const size_t row = (size_t)t * a.B + b;
            TrajIO io;
            io.field_in = fin; io.field_out = fout;
            io.p_in = a.p_in ? a.p_in + row * fs : nullptr;
            io.u_in = a.u_in ? a.u_in + row : nullptr;
            io.p_out = (a.p_out && t == nt - 1) ? a.p_out + (size_t)b * fs : nullptr;
            io.seed = a.seed; io.chain = a.chain0 + (uint64_t)b; io.traj = a.traj + (uint64_t)t;
            io.beta = a.beta; io.dt = a.dt; io.nstep = a.nstep;
            io.out_dH = a.s_out ? a.s_out + row : nullptr;
            io.out_expmdH = a.expmdH ? a.expmdH + row : nullptr;
            io.out_acc = a.acc ? a.acc + row : nullptr;
            io.out_plaq = a.plaq ? a.plaq + row : nullptr;
            io.out_Q = a.topo ? a.topo + row : nullptr;
            io.out_h0 = a.h0 ? a.h0 + row : nullptr;
            io.out_h1 = a.h1 ? a.h1 + row : nullptr;
            io.first = t == 0; io.last = t == nt - 1;
            if (a.mode == MODE_FT_HMC) ft_hmc_trajectory(en, io); else hmc_trajectory(en, io);
        }
    } break;
    default: break;
    }
}

// the plain-HMC programs only (no flow): what k_chain_plain runs.  Same code as the corresponding cases of run_chain, but
// a kernel that contains nothing else needs a fifth of the registers and several CTAs fit on an SM.
template <class E>
FT_HD void run_chain_plain(Engine<E>& en, const ChainArgs& a, int b) {
    E& ex = en.ex;
    const size_t fs = (size_t)2 * en.Vg;
    const double* fin = a.field_in + (size_t)b * fs;
    double* fout = a.field_out ? a.field_out + (size_t)b * fs : nullptr;
    ex.sync();
    if (a.mode == MODE_LEAPFROG) {
        en.load_field(en.oX, fin);
        en.for_links([&](int, int gi) { en.wsP[gi] = a.p_in[(size_t)b * fs + gi]; });
        ex.sync();
        if constexpr (E::kCluster) leapfrog_resident(en, a.dt, a.nstep, en.wsP, [&]() { en.wilson_force(a.beta, 1); });
        else leapfrog_plain_fused<E, true>(en, a.beta, a.dt, a.nstep, en.wsP);
        en.store_field(fout, en.oX);
        en.for_links([&](int, int gi) { a.p_out[(size_t)b * fs + gi] = en.wsP[gi]; });
        ex.sync();
        return;
    }
    const int nt = a.ntraj < 1 ? 1 : a.ntraj;
    for (int t = 0; t < nt; ++t) {
        const size_t row = (size_t)t * a.B + b;
        TrajIO io;
        io.field_in = fin; io.field_out = fout;
        io.p_in = a.p_in ? a.p_in + row * fs : nullptr;
        io.u_in = a.u_in ? a.u_in + row : nullptr;
        io.p_out = (a.p_out && t == nt - 1) ? a.p_out + (size_t)b * fs : nullptr;
        io.seed = a.seed; io.chain = a.chain0 + (uint64_t)b; io.traj = a.traj + (uint64_t)t;
        io.beta = a.beta; io.dt = a.dt; io.nstep = a.nstep;
        io.out_dH = a.s_out ? a.s_out + row : nullptr;
        io.out_expmdH = a.expmdH ? a.expmdH + row : nullptr;
        io.out_acc = a.acc ? a.acc + row : nullptr;
        io.out_plaq = a.plaq ? a.plaq + row : nullptr;
        io.out_Q = a.topo ? a.topo + row : nullptr;
        io.out_h0 = nullptr; io.out_h1 = nullptr;
        io.first = t == 0; io.last = t == nt - 1;
        if constexpr (E::kCluster) hmc_trajectory<E, true>(en, io); else hmc_trajectory_plain(en, io);
    }
}

}  // namespace fthmc
