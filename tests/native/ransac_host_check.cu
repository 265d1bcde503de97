// Test-only: runs the float64 building blocks of sfmfromscratch_b200/csrc/ransac_math.cuh on the
// HOST so that `-m "not gpu"` tests can compare them with the oracle where no GPU exists.  Not
// part of libsfmb200.so and never loaded by the product package.
#include <cstdint>
#include <cstring>

#include "../../sfmfromscratch_b200/csrc/ransac_math.cuh"

extern "C" {

// F [iters][9]; counts [iters]; cand [iters][48] and valid [iters] when pose != 0.
__attribute__((visibility("default")))
void ransac_host_eval(const double* p1, const double* p2, int n, const int32_t* samples, int iters, double thr,
                      int pose, const double* K1, const double* K2, const double* Rb, const double* Tb,
                      double* F_out, int32_t* counts, double* cand_out, uint32_t* valid) {
    double P1[12];
    if (pose) projection3x4(K1, Rb, Tb, P1);
    for (int it = 0; it < iters; ++it) {
        double x1[8], y1[8], x2[8], y2[8];
        for (int j = 0; j < 8; ++j) {
            const int i = samples[it * 8 + j];
            x1[j] = p1[2 * i]; y1[j] = p1[2 * i + 1]; x2[j] = p2[2 * i]; y2[j] = p2[2 * i + 1];
        }
        double* F = F_out + (size_t)it * 9;
        fundamental_8pt(x1, y1, x2, y2, F);
        if (pose) {
            double* cand = cand_out + (size_t)it * 48;
            pose_candidates(F, K1, K2, cand);
            valid[it] = 0;
            for (int c = 0; c < 4; ++c) {
                double P2[12];
                projection3x4(K2, cand + c * 12, cand + c * 12 + 9, P2);
                bool ok = true;
                for (int i = 0; i < n && ok; ++i)
                    ok = point_in_front(P1, P2, Rb, Tb, cand + c * 12, cand + c * 12 + 9,
                                        make_double2(p1[2 * i], p1[2 * i + 1]), make_double2(p2[2 * i], p2[2 * i + 1]));
                if (ok) valid[it] |= 1u << c;
            }
        }
        int cnt = 0;
        for (int i = 0; i < n; ++i)
            cnt += is_inlier(F, make_double2(p1[2 * i], p1[2 * i + 1]), make_double2(p2[2 * i], p2[2 * i + 1]), thr) ? 1 : 0;
        counts[it] = cnt;
    }
}

}
