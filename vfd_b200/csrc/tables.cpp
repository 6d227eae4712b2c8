// tables.cpp — host-side construction of the lookup tables the kernels read.
//  * KernelTables::build reproduces PrecomputedDFSPHCubicKernel::SetRadius
//    (reference: Kernel/DFSPHKernels.h:13-41, CalculateW :88-107, CalculateGradientW :113-134)
//    with the same fp32 operations in the same order, so the tables are bit-identical, and then
//    precombines the midpoint rule of GetW/GetGradientW (:54-80) into one table per function.
//  * build_halton_table regenerates the 16 384 unit-sphere Halton points of the reference's data file
//    (HaltonVec323.cuh:4-1643; used at DFSPHKernels.cu:917-924): point i = (s cos phi, s sin phi, z),
//    z = 1 - 2 H3(i), phi = 2 pi H2(i), s = sqrt(1 - z^2)  (SURVEY.md F12).
// Compile without FMA contraction / fast-math: this file is pure fp32/fp64 host arithmetic.
#include "solver.h"
#include <cmath>

namespace vfd {

static float calc_w(float r, float radius, float k) {
    float res = 0.0f;
    const float q = r / radius;
    if (q <= 1.0f) {
        if (q <= 0.5f) {
            const float q2 = q * q;
            const float q3 = q2 * q;
            res = k * (6.0f * q3 - 6.0f * q2 + 1.0f);
        } else {
            const float q1 = 1.0f - q;
            const float q3 = q1 * q1 * q1;
            res = k * (2.0f * q3);
        }
    }
    return res;
}

// x component of grad W at (x, 0, 0)
static float calc_grad_w_x(float x, float radius, float l) {
    const float rl = std::sqrt((x * x + 0.0f * 0.0f) + 0.0f * 0.0f);
    const float q = rl / radius;
    float res = 0.0f;
    if (rl > 1.0e-5f && (q <= 1.0f)) {
        const float gqx = x * (1.0f / (rl * radius));
        if (q <= 0.5f) {
            res = l * q * (3.0f * q - 2.0f) * gqx;
        } else {
            const float factor = 1.0f - q;
            res = l * (-factor * factor) * gqx;
        }
    }
    return res;
}

void KernelTables::build(float rad) {
    const float pi = static_cast<float>(3.14159265358979323846);
    radius = rad;
    radius2 = radius * radius;
    const float radius3 = radius * radius * radius;
    k = 8.0f / (pi * radius3);
    l = 48.0f / (pi * radius3);
    wZero = calc_w(std::sqrt(0.0f), radius, k);
    const float stepSize = radius / static_cast<float>(VFD_LUT_RES - 1);
    invStep = 1.0f / stepSize;
    W.assign(VFD_LUT_RES, 0.0f);
    gradW.assign(VFD_LUT_RES + 1, 0.0f);
    for (unsigned int i = 0; i < VFD_LUT_RES; i++) {
        const float posX = stepSize * static_cast<float>(i);
        W[i] = calc_w(posX, radius, k);
        if (posX > 1.0e-9f) gradW[i] = calc_grad_w_x(posX, radius, l) / posX;
        else gradW[i] = 0.0f;
    }
    gradW[VFD_LUT_RES] = 0.0f;
    Wc.assign(VFD_LUT_RES, 0.0f);
    Gc.assign(VFD_LUT_RES, 0.0f);
    for (unsigned int p = 0; p + 1 < VFD_LUT_RES; p++) {
        Wc[p] = 0.5f * (W[p] + W[p + 1]);
        Gc[p] = 0.5f * (gradW[p] + gradW[p + 1]);
    }
}

static double radical_inverse(unsigned int i, unsigned int base) {
    double f = 1.0, r = 0.0;
    while (i > 0) {
        f /= (double)base;
        r += f * (double)(i % base);
        i /= base;
    }
    return r;
}

void build_halton_table(std::vector<float>& out) {
    const double PI = 3.14159265358979323846;
    out.resize(VFD_HALTON_N);
    for (unsigned int i = 0; i < VFD_HALTON_N / 3u; i++) {
        const double z = 1.0 - 2.0 * radical_inverse(i, 3);
        const double phi = 2.0 * PI * radical_inverse(i, 2);
        const double s = std::sqrt(1.0 - z * z);
        const double v[3] = { s * std::cos(phi), s * std::sin(phi), z };
        // the data file holds exact zeros where double arithmetic leaves +-1e-16 (sin(pi), cos(pi/2) ...)
        for (int k = 0; k < 3; k++) out[3 * i + k] = std::fabs(v[k]) < 1.0e-15 ? 0.0f : (float)v[k];
    }
}

} // namespace vfd
