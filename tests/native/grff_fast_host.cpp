// Host build of the FP32 voxel evaluation (raytracinggrff_b200/csrc/grff_fast.cuh) for the CPU tests: the
// very source the kernels compile, with libm standing in for the MUFU approximations.
#include <stdint.h>
#include "../../raytracinggrff_b200/csrc/grff_fast.cuh"

using namespace rtgrff;

extern "C" {

// fc: {c_su, cv_hi, cv_lo, lnl_cold, lnl_hot, kff, srcc}; vox: n x {dz, T, ne, B, cth, sth, scale}; out: n x {aL,aR,bL,bR}
void grff_fast_eval(const float *fc, const float *vox, int64_t n, int ff_on, float *out, uint8_t *ok)
{
    FreqCF f;
    f.c_su = fc[0]; f.cv_hi = fc[1]; f.cv_lo = fc[2]; f.lnl_cold = fc[3]; f.lnl_hot = fc[4]; f.kff = fc[5]; f.srcc = fc[6];
    for (int64_t i = 0; i < n; ++i) {
        const float *v = vox + i * 7;
        const FastOp o = voxel_op_f32(f, v[0], v[1], v[2], v[3], v[4], v[5], v[6], ff_on != 0);
        out[i * 4 + 0] = o.aL; out[i * 4 + 1] = o.aR; out[i * 4 + 2] = o.bL; out[i * 4 + 3] = o.bR;
        ok[i] = o.ok ? 1 : 0;
    }
}
}
