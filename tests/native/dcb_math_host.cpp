// TEST INFRASTRUCTURE: deepcomp_b200/csrc/dcb_math.cuh -- the table-driven fp64 math of the step kernels -- compiled for the
// HOST with g++, so that its accuracy claims are checked by the CPU test suite (tests/test_device_math_host.py) on the very
// source the kernels are built from.  Only the device intrinsics the header uses are supplied here.
#include <cmath>
#include <cstdint>
#include <cstring>

static inline int __double2hiint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int __double2loint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x, &u, 8); return x;
}

#include "dcb_math.cuh"

extern "C" {

// tables: the 80 doubles of MathTables (inv, l2c, ex2, pwm, pwe), pw: the 10 binomial coefficients -- built by the caller the
// way dcb_create builds them (dcb_api.cu: host_math_tables, DevParams::pw)
void mh_log2(const double *tables, const double *x, double *out, int n) {
    const MathTables *t = reinterpret_cast<const MathTables *>(tables);
    for (int i = 0; i < n; i++) out[i] = dcb_log2(t, x[i]);
}
void mh_exp2(const double *tables, const double *y, double *out, int n) {
    const MathTables *t = reinterpret_cast<const MathTables *>(tables);
    for (int i = 0; i < n; i++) out[i] = dcb_exp2(t, y[i]);
}
void mh_log2_1p(const double *tables, const double *s, double *out, int n) {
    const MathTables *t = reinterpret_cast<const MathTables *>(tables);
    for (int i = 0; i < n; i++) out[i] = dcb_log2_1p(t, s[i]);
}
void mh_rcp(const double *x, double *out, int n) {
    for (int i = 0; i < n; i++) out[i] = dcb_rcp(x[i]);
}
// snr of squared distances in [1, 2^32): dcb_snr_inrange takes the binary exponent modulo 16; exponents 16..31 are its
// value times k16 = 2^(-16 h) (the far pairs of the interference pass, dcb_wide.cu: snr_of_d2_anywhere)
void mh_snr(const double *tables, const double *pw, double h, const double *d2, double *out, int n) {
    const MathTables *t = reinterpret_cast<const MathTables *>(tables);
    const double k16 = dcb_exp2(t, -16.0 * h);
    for (int i = 0; i < n; i++) {
        const double s = dcb_snr_inrange(t, pw, d2[i]);
        out[i] = d2[i] < 65536.0 ? s : s * k16;
    }
}

}  // extern "C"
