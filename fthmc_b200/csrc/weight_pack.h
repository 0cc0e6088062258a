// weight_pack.h -- host-side re-layout of one coupling layer's CNN weights into the packed block the
// chain engine stages in shared memory (offsets in chain_engine.cuh).
//
// Input ("raw") is the reference's own parameter order for layer.plaq_coupling.net
// (ipynb/field_transformation.py:84-99): conv0.weight (8,2,3,3), conv0.bias (8), conv1.weight (8,8,3,3),
// conv1.bias (8), conv2.weight (3,8,3,3), conv2.bias (3)  == 955 doubles, row-major.
//
// Canonical orientation: the engine always sees stripes running along "rows" r and the 4-periodic
// mask pattern across "columns" c.  For mask_mu == 0 (r=n0, c=n1) the kernel is used as is; for
// mask_mu == 1 (r=n1, c=n0) the two spatial kernel indices are swapped.  The cyclic column shift by
// mask_off commutes with a circular convolution and needs no weight change.
#pragma once
#include "chain_engine.cuh"

namespace fthmc {

constexpr int RAW_DOUBLES = 8 * 2 * 9 + 8 + 8 * 8 * 9 + 8 + 3 * 8 * 9 + 3;   // 955

inline void pack_layer(const double* raw, int mu, double* out) {
    const double* w1 = raw;                 // [o8][ci2][i][j]
    const double* b1 = w1 + 144;
    const double* w2 = b1 + 8;              // [o8][ci8][i][j]
    const double* b2 = w2 + 576;
    const double* w3 = b2 + 8;              // [o3][ci8][i][j]
    const double* b3 = w3 + 216;
    auto W1 = [&](int o, int ci, int a, int b) { return mu == 0 ? w1[((o * 2 + ci) * 3 + a) * 3 + b] : w1[((o * 2 + ci) * 3 + b) * 3 + a]; };
    auto W2 = [&](int o, int ci, int a, int b) { return mu == 0 ? w2[((o * 8 + ci) * 3 + a) * 3 + b] : w2[((o * 8 + ci) * 3 + b) * 3 + a]; };
    auto W3 = [&](int o, int ci, int a, int b) { return mu == 0 ? w3[((o * 8 + ci) * 3 + a) * 3 + b] : w3[((o * 8 + ci) * 3 + b) * 3 + a]; };
    for (int i = 0; i < PACK_DOUBLES; ++i) out[i] = 0.0;
    for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a) for (int ci = 0; ci < 2; ++ci) for (int o = 0; o < 8; ++o)
        out[OFF_W1F + ((b * 3 + a) * 2 + ci) * 8 + o] = W1(o, ci, a, b);
    // column classes after the shift by mask_off: 0 active, 1 and 2 frozen, 3 passive.  Non-frozen
    // plaquettes enter the CNN as (cos 0, sin 0) = (1, 0): their cos-channel taps fold into the bias.
    for (int q = 0; q < 4; ++q) for (int o = 0; o < 8; ++o) {
        double acc = b1[o];
        for (int b = 0; b < 3; ++b) {
            int cls = ((q + b - 1) % 4 + 4) % 4;
            if (cls == 1 || cls == 2) continue;
            for (int a = 0; a < 3; ++a) acc += W1(o, 0, a, b);
        }
        out[OFF_B1 + q * 8 + o] = acc;
    }
    for (int ci = 0; ci < 8; ++ci) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int o = 0; o < 8; ++o) {
        out[OFF_W2F + ((ci * 3 + a) * 3 + b) * 8 + o] = W2(o, ci, a, b);
        out[OFF_W2T + ((o * 3 + a) * 3 + b) * 8 + ci] = W2(o, ci, a, b);
    }
    for (int o = 0; o < 8; ++o) out[OFF_B2 + o] = b2[o];
    for (int ci = 0; ci < 8; ++ci) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int o = 0; o < 3; ++o) {
        out[OFF_W3F + ((ci * 3 + a) * 3 + b) * 4 + o] = W3(o, ci, a, b);
        out[OFF_W3T + ((o * 3 + a) * 3 + b) * 8 + ci] = W3(o, ci, a, b);
    }
    for (int o = 0; o < 3; ++o) out[OFF_B3 + o] = b3[o];
    for (int o = 0; o < 8; ++o) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int ci = 0; ci < 2; ++ci)
        out[OFF_W1T + ((o * 3 + a) * 3 + b) * 2 + ci] = W1(o, ci, a, b);
}

// Adjoint of the forward half of pack_layer: canonical gradient block (GRAD_DOUBLES doubles: conv1 [b][a][ci][o], the
// column-class sums S_q[o] in the conv1-bias slots, conv2 [ci][a][b][o], bias2, conv3 [ci][a][b][4], bias3) -> gradient
// with respect to the raw parameters in the reference's order (955 doubles).  The packed conv1 bias of class q is
// b1[o] + sum_{b: class(q+b-1) not frozen} sum_a W1[o][0][a][b], hence db1[o] = sum_q S_q[o] and every W1[o][0][a][b]
// collects S_q[o] of the classes q whose kernel column b looks at a non-frozen column.
inline void unpack_grad_layer(const double* cg, int mu, double* raw) {
    double* w1 = raw;
    double* b1 = w1 + 144;
    double* w2 = b1 + 8;
    double* b2 = w2 + 576;
    double* w3 = b2 + 8;
    double* b3 = w3 + 216;
    for (int i = 0; i < RAW_DOUBLES; ++i) raw[i] = 0.0;
    auto I1 = [&](int o, int ci, int a, int b) { return mu == 0 ? ((o * 2 + ci) * 3 + a) * 3 + b : ((o * 2 + ci) * 3 + b) * 3 + a; };
    auto I2 = [&](int o, int ci, int a, int b) { return mu == 0 ? ((o * 8 + ci) * 3 + a) * 3 + b : ((o * 8 + ci) * 3 + b) * 3 + a; };
    for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a) for (int ci = 0; ci < 2; ++ci) for (int o = 0; o < 8; ++o)
        w1[I1(o, ci, a, b)] += cg[OFF_W1F + ((b * 3 + a) * 2 + ci) * 8 + o];
    for (int q = 0; q < 4; ++q) for (int o = 0; o < 8; ++o) {
        const double sq = cg[OFF_B1 + q * 8 + o];
        b1[o] += sq;
        for (int b = 0; b < 3; ++b) {
            int cls = ((q + b - 1) % 4 + 4) % 4;
            if (cls == 1 || cls == 2) continue;
            for (int a = 0; a < 3; ++a) w1[I1(o, 0, a, b)] += sq;
        }
    }
    for (int ci = 0; ci < 8; ++ci) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int o = 0; o < 8; ++o)
        w2[I2(o, ci, a, b)] += cg[OFF_W2F + ((ci * 3 + a) * 3 + b) * 8 + o];
    for (int o = 0; o < 8; ++o) b2[o] += cg[OFF_B2 + o];
    for (int ci = 0; ci < 8; ++ci) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int o = 0; o < 3; ++o)
        w3[I2(o, ci, a, b)] += cg[OFF_W3F + ((ci * 3 + a) * 3 + b) * 4 + o];
    for (int o = 0; o < 3; ++o) b3[o] += cg[OFF_B3 + o];
}

}  // namespace fthmc
