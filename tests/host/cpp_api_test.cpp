// Drives the deconvolution loop through the C++ host mirror (include/mvdecon.hpp) exactly like code written against the reference's
// classes would: DeconView / DeconViews / PsiInitFromRAI / MultiViewDeconvolutionSeq.  Inputs come as raw float32 files written by the
// Python test (tests/test_cpp_api.py), the result goes back as raw files and is compared with the oracle there.  TEST ONLY.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include "mvdecon.hpp"

using namespace mvrecon;

static std::vector<float> read_f32(const std::string& path, long long n) {
    std::vector<float> v((size_t)n);
    std::ifstream f(path, std::ios::binary);
    if (!f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * 4))) throw Error("cannot read " + path);
    return v;
}
static void write_f32(const std::string& path, const std::vector<float>& v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * 4));
}

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "usage: cpp_api_test <dir>\n"; return 2; }
    const std::string dir = argv[1];
    try {
        int nx, ny, nz, V, kx, ky, kz, ptype, iters;
        float lambda;
        {
            std::ifstream m(dir + "/meta.txt");
            if (!(m >> nx >> ny >> nz >> V >> kx >> ky >> kz >> ptype >> lambda >> iters)) throw Error("bad meta.txt");
        }
        const Dims dims{nx, ny, nz}, kd{kx, ky, kz};
        std::vector<std::vector<float>> img, w, psf;
        for (int v = 0; v < V; ++v) {
            img.push_back(read_f32(dir + "/img" + std::to_string(v) + ".f32", numElements(dims)));
            w.push_back(read_f32(dir + "/w" + std::to_string(v) + ".f32", numElements(dims)));
            psf.push_back(read_f32(dir + "/psf" + std::to_string(v) + ".f32", numElements(kd)));
        }
        const std::vector<float> psi0 = read_f32(dir + "/psi0.f32", numElements(dims));
        const std::vector<float> mx = read_f32(dir + "/max.f32", V);

        std::vector<DeconView> list;
        for (int v = 0; v < V; ++v) list.emplace_back(Img(img[v], dims), Img(w[v], dims), Img(psf[v], kd), (PSFTYPE)ptype, "view " + std::to_string(v));
        DeconViews views(std::move(list), 0, lambda);
        for (int v = 0; v < V; ++v) {
            write_f32(dir + "/k1_" + std::to_string(v) + ".f32", views.getViews()[v].getPSF().getKernel1());
            write_f32(dir + "/k2_" + std::to_string(v) + ".f32", views.getViews()[v].getPSF().getKernel2());
        }
        PsiInitFromRAI init(Img(psi0, dims), mx);
        MultiViewDeconvolutionSeq decon(views, iters, init);
        if (!decon.initWasSuccessful()) throw Error("init failed");
        std::ofstream st(dir + "/stats.txt");
        st.precision(17);
        while (decon.currentIteration() < iters)
            for (const IterationStatistics& s : decon.runNextIteration()) st << s.sumChange << " " << s.maxChange << "\n";
        write_f32(dir + "/psi_out.f32", decon.getPSI());

        // error behaviour: mismatching view sizes are rejected like DeconViews.java:61-64
        bool threw = false;
        try {
            std::vector<float> small(8, 1.f);
            std::vector<DeconView> bad;
            bad.emplace_back(Img(img[0], dims), Img(w[0], dims), Img(psf[0], kd));
            bad.emplace_back(Img(small, Dims{2, 2, 2}), Img(small, Dims{2, 2, 2}), Img(psf[0], kd));
            DeconViews nope(std::move(bad));
        } catch (const Error&) { threw = true; }
        if (!threw) throw Error("mismatching view dimensions were accepted");

        // operator level: one block through ComputeBlockSeqThreadB200 == the same update on the whole (small) volume as a single block
        ComputeBlockSeqThreadB200Factory factory(MultiViewDeconvolution::minValue, lambda, dims, {0});
        ComputeBlockSeqThreadB200 worker = factory.create(0);
        worker.getPsiBlockTmp() = psi0;
        const IterationStatistics bs = worker.runIteration(img[0].data(), w[0].data(), mx[0], views.getViews()[0].getPSF());
        write_f32(dir + "/block_out.f32", worker.getPsiBlockTmp());
        st << bs.sumChange << " " << bs.maxChange << "\n";

        // block-wise driver: halo'd blocks of 16^3 through the operator, delayed paste-back, one iteration from psi0
        std::vector<float> psi = psi0;
        const Dims bsz{16, 16, 16};
        ComputeBlockSeqThreadB200Factory bf(MultiViewDeconvolution::minValue, lambda, bsz, {0});
        const Dims k1d = views.getViews()[0].getPSF().getKernel1Dims();
        const std::vector<Block> blocks = divideIntoBlocks(dims, bsz, Dims{2 * k1d[0] - 1, 2 * k1d[1] - 1, 2 * k1d[2] - 1});
        if (blocks.size() < 8) throw Error("expected a multi-block plan");
        size_t covered = 0;
        for (const Block& b : blocks) covered += (size_t)numElements(b.effectiveSize);
        if ((long long)covered != numElements(dims)) throw Error("effective regions do not tile the volume");
        for (const IterationStatistics& s : runNextIterationBlocked(psi.data(), dims, views, mx, bf, bsz)) st << s.sumChange << " " << s.maxChange << "\n";
        write_f32(dir + "/psi_blocked.f32", psi);
    } catch (const std::exception& e) {
        std::cerr << "cpp_api_test: " << e.what() << "\n";
        return 1;
    }
    std::cout << "ok\n";
    return 0;
}
