// Host emulation test of the FFT codelets / column pass bodies (no GPU needed).
// Build: g++ -std=c++17 -O1 -DMVD_HOST_EMU -I multiview-reconstruction_b200/csrc tests/host/fft_emu_test.cpp
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

#include "backend.h"

using namespace mvd;
typedef std::complex<double> cd;

static std::vector<cd> naive_dft(const std::vector<cd>& x) {
    const int n = (int)x.size();
    std::vector<cd> X(n);
    for (int k = 0; k < n; ++k) {
        cd s = 0;
        for (int j = 0; j < n; ++j) s += x[j] * std::polar(1.0, -2.0 * M_PI * double((long long)j * k % n) / n);
        X[k] = s;
    }
    return X;
}

template <class P>
static int test_plan() {
    constexpr int N = P::N, W = P::W;
    const int nx = W + 3, nb = 2;                       // 2 column groups (second one ragged), 2 batch lines
    const long long stride_n = nx + 1, stride_b = (long long)N * stride_n + 5;
    std::vector<cpx> data(nb * stride_b + 16), orig;
    std::mt19937 rng(N);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    for (auto& c : data) c = cpx{U(rng), U(rng)};
    orig = data;
    std::vector<cpx> tw(StageTw<P>::NTW + 1);
    fill_stage_tw<P>(tw.data());
    std::vector<cpx> khat(data.size());
    for (auto& c : khat) c = cpx{U(rng), U(rng)};

    ColArgs a{data.data(), khat.data(), tw.data(), stride_n, stride_b, nx, 0, 0, 0};
    std::vector<cpx> sm(ColSmem<P>::bytes() / sizeof(cpx) + 1);
    HostExec ex(P::THREADS);
    const int gx = (nx + W - 1) / W;
    auto run = [&](auto modec) {
        constexpr int MODE = decltype(modec)::value;
        for (int by = 0; by < nb; ++by)
            for (int bx = 0; bx < gx; ++bx) col_pass_body<P, MODE>(ex, a, bx, by, sm.data());
    };
    int fails = 0;
    // forward vs naive DFT
    run(std::integral_constant<int, COL_FWD>{});
    double maxerr = 0, maxref = 0;
    for (int b = 0; b < nb; ++b)
        for (int x = 0; x < nx; ++x) {
            std::vector<cd> in(N);
            for (int n = 0; n < N; ++n) { cpx c = orig[b * stride_b + n * stride_n + x]; in[n] = cd(c.x, c.y); }
            auto X = naive_dft(in);
            for (int n = 0; n < N; ++n) {
                cpx c = data[b * stride_b + n * stride_n + x];
                cd ref = X[plan_freq_of_pos<P>(n)];
                maxerr = std::max(maxerr, std::abs(cd(c.x, c.y) - ref));
                maxref = std::max(maxref, std::abs(ref));
            }
        }
    if (!(maxerr < 2e-5 * maxref * std::log2((double)N))) { ++fails; }
    std::printf("N=%4d (%d,%d,%d) T=%d W=%d  fwd err %.3g (ref %.3g)", N, P::R1, P::R2, P::R3, P::T, W, maxerr, maxref);
    // untouched padding elements must stay untouched
    std::vector<cpx> fwd = data;
    // inverse returns N * x
    run(std::integral_constant<int, COL_INV>{});
    double ierr = 0;
    for (int b = 0; b < nb; ++b)
        for (int x = 0; x < nx; ++x)
            for (int n = 0; n < N; ++n) {
                cpx c = data[b * stride_b + n * stride_n + x], o = orig[b * stride_b + n * stride_n + x];
                ierr = std::max(ierr, (double)std::hypot(c.x / N - o.x, c.y / N - o.y));
            }
    if (!(ierr < 1e-5)) ++fails;
    std::printf("  inv err %.3g", ierr);
    // elements outside the addressed columns untouched
    for (int b = 0; b < nb; ++b)
        for (int n = 0; n < N; ++n) {
            cpx c = data[b * stride_b + n * stride_n + nx], o = orig[b * stride_b + n * stride_n + nx];
            if (c.x != o.x || c.y != o.y) { ++fails; std::printf(" [padding touched]"); b = nb; break; }
        }
    // fused conv == fwd, multiply, inv
    data = orig;
    run(std::integral_constant<int, COL_CONV>{});
    std::vector<cpx> fused = data;
    data = fwd;
    for (int b = 0; b < nb; ++b)
        for (int x = 0; x < nx; ++x)
            for (int n = 0; n < N; ++n) {
                long long i = b * stride_b + n * stride_n + x;
                data[i] = cmul(data[i], khat[i]);
            }
    run(std::integral_constant<int, COL_INV>{});
    double cerr = 0, cref = 0;
    for (int b = 0; b < nb; ++b)
        for (int x = 0; x < nx; ++x)
            for (int n = 0; n < N; ++n) {
                long long i = b * stride_b + n * stride_n + x;
                cerr = std::max(cerr, (double)std::hypot(fused[i].x - data[i].x, fused[i].y - data[i].y));
                cref = std::max(cref, (double)std::hypot(data[i].x, data[i].y));
            }
    if (!(cerr < 1e-5 * cref)) ++fails;
    std::printf("  conv err %.3g (ref %.3g) %s\n", cerr, cref, fails ? "FAIL" : "ok");
    return fails;
}

int main() {
    int fails = 0;
#define TP(N, R1, R2, R3, T, W) fails += test_plan<Plan<N, R1, R2, R3, T, W>>();
    MVD_TEST_PLANS(TP)
    std::printf("%s\n", fails ? "FAILED" : "ALL OK");
    return fails ? 1 : 0;
}
