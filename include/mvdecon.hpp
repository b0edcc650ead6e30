// C++ host-side mirror of the reference's operator interface for the deconvolution path, header-only, on top of the C ABI
// (include/mvdecon.h).  The reference is compiled (Java) code whose toolchain is absent from the build image, so this is the
// compiled-language host side a caller links against; class and method names follow net.preibisch.mvrecon.process.deconvolution
// so that code written against the reference reads the same:
//
//   PSFTYPE, DeconViewPSF            M/process/deconvolution/DeconViewPSF.java:52,119-254
//   DeconView, DeconViews            M/process/deconvolution/DeconView.java:118-184, DeconViews.java:44-81
//   PsiInit*                         M/process/deconvolution/init/PsiInit*.java
//   MultiViewDeconvolution(Seq|Mul)  M/process/deconvolution/MultiViewDeconvolution.java:90-200, ...Seq.java:58-180, ...Mul.java:116-245
//   ComputeBlockSeqThreadB200(Factory)  .../iteration/sequential/ComputeBlockSeqThread.java:54-61, ComputeBlockSeqThreadCUDAFactory.java:41-64
//   IterationStatistics              .../iteration/IterationStatistics.java
//
// Volumes are float32, x fastest; dimension triples are (x, y, z) like the reference's Dimensions / Block.  Errors of the library
// surface as mvrecon::Error carrying mvd_last_error().  Nothing here computes: all arithmetic is inside libmvdecon.so.
#ifndef MVDECON_HPP
#define MVDECON_HPP

#include <algorithm>
#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mvdecon.h"

namespace mvrecon {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc) {
    if (rc != 0) throw Error(mvd_last_error());
}

using Dims = std::array<int, 3>;   // (x, y, z)
inline long long numElements(const Dims& d) { return (long long)d[0] * d[1] * d[2]; }

// a caller-owned float volume (no copy)
struct Img {
    const float* data = nullptr;
    Dims dims{0, 0, 0};
    Img() = default;
    Img(const float* p, const Dims& d) : data(p), dims(d) {}
    Img(const std::vector<float>& v, const Dims& d) : data(v.data()), dims(d) {
        if ((long long)v.size() != numElements(d)) throw Error("Img: size does not match the dimensions");
    }
};

enum class PSFTYPE : int { OPTIMIZATION_II = MVD_PSF_OPTIMIZATION_II, OPTIMIZATION_I = MVD_PSF_OPTIMIZATION_I,
                           EFFICIENT_BAYESIAN = MVD_PSF_EFFICIENT_BAYESIAN, INDEPENDENT = MVD_PSF_INDEPENDENT };

struct IterationStatistics {
    double sumChange = 0, maxChange = -1;
};

class DeconViewPSF {
  public:
    DeconViewPSF(Img psf, PSFTYPE type) : psf_(psf), type_(type) {}
    PSFTYPE getPSFType() const { return type_; }
    const Img& getPSF() const { return psf_; }
    // available after DeconViews has initialised the views (DeconViewPSF.init)
    const std::vector<float>& getKernel1() const { return k1_; }
    const std::vector<float>& getKernel2() const { return k2_; }
    const Dims& getKernel1Dims() const { return k1d_; }
    const Dims& getKernel2Dims() const { return k2d_; }

  private:
    friend class DeconViews;
    Img psf_;
    PSFTYPE type_;
    std::vector<float> k1_, k2_;
    Dims k1d_{0, 0, 0}, k2d_{0, 0, 0};
};

class DeconView {
  public:
    // weight.data == nullptr: the weight is generated on the device later (DeconViews::makeBlendingWeights + normalizeWeights)
    DeconView(Img image, Img weight, Img kernel, PSFTYPE psfType = PSFTYPE::INDEPENDENT, std::string title = "")
        : image_(image), weight_(weight), psf_(kernel, psfType), title_(std::move(title)) {
        if (weight.data && weight.dims != image.dims) throw Error("image and weight must be volumes of identical size");
    }
    const Img& getImage() const { return image_; }
    const Img& getWeight() const { return weight_; }
    DeconViewPSF& getPSF() { return psf_; }
    const DeconViewPSF& getPSF() const { return psf_; }
    const std::string& getTitle() const { return title_; }

  private:
    Img image_, weight_;
    DeconViewPSF psf_;
    std::string title_;
};

// DeconViews: dimension check, PSF init in list order; owns the resident device context
class DeconViews {
  public:
    DeconViews(std::vector<DeconView> views, int device = 0, float lambda = 0.f, float minValue = 1e-4f, int normQuirkThreads = 0,
               int maxFftLen = 0)
        : views_(std::move(views)) {
        if (views_.empty()) throw Error("no views");
        dims_ = views_[0].getImage().dims;
        for (const DeconView& v : views_) {                       // DeconViews.java:61-64
            if (v.getImage().dims != dims_) throw Error("dimensions of all views must be identical");
            if (v.getPSF().getPSFType() != views_[0].getPSF().getPSFType()) throw Error("all views must use the same PSFTYPE");
        }
        mvd_config cfg{};
        cfg.device = device;
        for (int d = 0; d < 3; ++d) cfg.dims[d] = dims_[d];
        cfg.num_views = (int)views_.size();
        cfg.psf_type = (int)views_[0].getPSF().getPSFType();
        cfg.lambda = lambda;
        cfg.min_value = minValue;
        cfg.shard_lo = 0; cfg.shard_hi = dims_[2]; cfg.local_z0 = 0; cfg.local_nz = dims_[2];
        cfg.max_fft_len = maxFftLen;
        cfg.norm_quirk_threads = normQuirkThreads;
        mvd_context* c = nullptr;
        check(mvd_create(&cfg, &c));
        ctx_.reset(c, [](mvd_context* p) { mvd_destroy(p); });
        for (size_t i = 0; i < views_.size(); ++i) {
            const DeconView& v = views_[i];
            check(mvd_set_view(c, (int)i, v.getImage().data, v.getWeight().data));
            check(mvd_set_psf(c, (int)i, v.getPSF().getPSF().data, v.getPSF().getPSF().dims.data()));
        }
        check(mvd_init_views(c));                                  // psf.init( this, blockSize ) for every view
        for (size_t i = 0; i < views_.size(); ++i) {
            DeconViewPSF& p = views_[i].getPSF();
            check(mvd_get_kernel_dims(c, (int)i, 1, p.k1d_.data()));
            check(mvd_get_kernel_dims(c, (int)i, 2, p.k2d_.data()));
            p.k1_.resize((size_t)numElements(p.k1d_));
            p.k2_.resize((size_t)numElements(p.k2d_));
            check(mvd_get_kernel(c, (int)i, 1, p.k1_.data()));
            check(mvd_get_kernel(c, (int)i, 2, p.k2_.data()));
        }
    }
    std::vector<DeconView>& getViews() { return views_; }
    const Dims& getPSIDimensions() const { return dims_; }
    mvd_context* context() const { return ctx_.get(); }

    // weight masks on the device (BlendingRealRandomAccess / NormalizingRandomAccess)
    void makeBlendingWeights(int v, const Dims& boxMin, const Dims& boxMax, const std::array<float, 3>& border, const std::array<float, 3>& blending) {
        check(mvd_make_blending_weights(ctx_.get(), v, boxMin.data(), boxMax.data(), border.data(), blending.data()));
    }
    void normalizeWeights(double osemSpeedup = 1.0, bool additionalSmoothBlending = false, float maxDiffRange = 0.1f, float scalingRange = 0.05f) {
        check(mvd_normalize_weights(ctx_.get(), osemSpeedup, additionalSmoothBlending ? 1 : 0, maxDiffRange, scalingRange));
    }
    std::vector<float> getWeight(int v) const {
        std::vector<float> w((size_t)numElements(dims_));
        check(mvd_get_weight(ctx_.get(), v, w.data()));
        return w;
    }

  private:
    std::vector<DeconView> views_;
    Dims dims_{0, 0, 0};
    std::shared_ptr<mvd_context> ctx_;
};

// ---- PsiInit family (M/process/deconvolution/init/PsiInit.java) -------------------------------------------------------------------------
class PsiInit {
  public:
    virtual ~PsiInit() = default;
    virtual bool runInitialization(DeconViews& views) = 0;        // sets psi inside the context
    double getAvg() const { return avg_; }
    const std::vector<float>& getMax() const { return max_; }

  protected:
    double avg_ = 0;
    std::vector<float> max_;
};
class PsiInitFromRAI : public PsiInit {                            // PsiInitFromRAI.java: psi and the per-view maxima are given
  public:
    PsiInitFromRAI(Img psi, std::vector<float> max, double avg = 0) : psi_(psi) { max_ = std::move(max); avg_ = avg; }
    bool runInitialization(DeconViews& views) override {
        if (psi_.dims != views.getPSIDimensions() || max_.size() != views.getViews().size()) throw Error("PsiInitFromRAI: wrong dimensions");
        check(mvd_set_psi(views.context(), psi_.data));
        check(mvd_set_max_intensities(views.context(), max_.data()));
        return true;
    }

  private:
    Img psi_;
};
class PsiInitDevice : public PsiInit {
  public:
    PsiInitDevice(int type, double sigma) : type_(type), sigma_(sigma) {}
    bool runInitialization(DeconViews& views) override {
        max_.assign(views.getViews().size(), 0.f);
        check(mvd_psi_init(views.context(), type_, sigma_, &avg_, max_.data()));
        return true;
    }

  private:
    int type_;
    double sigma_;
};
struct PsiInitBlurredFused : PsiInitDevice { explicit PsiInitBlurredFused(double sigma = 5.0) : PsiInitDevice(MVD_PSI_FUSED_BLURRED, sigma) {} };
struct PsiInitAvgPrecise : PsiInitDevice { PsiInitAvgPrecise() : PsiInitDevice(MVD_PSI_AVG, 0) {} };
struct PsiInitAvgApprox : PsiInitDevice { PsiInitAvgApprox() : PsiInitDevice(MVD_PSI_APPROX_AVG, 0) {} };

// ---- MultiViewDeconvolution (MultiViewDeconvolution.java:90-200) --------------------------------------------------------------------------
class MultiViewDeconvolution {
  public:
    static constexpr float minValue = 1e-4f, minValueImg = 1.f, outsideValueImg = 0.f;     // MultiViewDeconvolution.java:48-50
    MultiViewDeconvolution(DeconViews& views, int numIterations, PsiInit& psiInit) : views_(views), numIterations_(numIterations) {
        initOk_ = psiInit.runInitialization(views);
        max_ = psiInit.getMax();
    }
    virtual ~MultiViewDeconvolution() = default;
    bool initWasSuccessful() const { return initOk_; }
    int currentIteration() const { return it_; }
    virtual std::vector<IterationStatistics> runNextIteration() = 0;
    void runIterations() {                                          // MultiViewDeconvolution.java:144-200
        while (it_ < numIterations_) runNextIteration();
    }
    void getPSI(float* out) const { check(mvd_get_psi(views_.context(), out)); }
    std::vector<float> getPSI() const {
        std::vector<float> psi((size_t)numElements(views_.getPSIDimensions()));
        getPSI(psi.data());
        return psi;
    }

  protected:
    DeconViews& views_;
    int numIterations_, it_ = 0;
    bool initOk_ = false;
    std::vector<float> max_;
};
class MultiViewDeconvolutionSeq : public MultiViewDeconvolution {   // OSEM: psi updated after every view
  public:
    using MultiViewDeconvolution::MultiViewDeconvolution;
    std::vector<IterationStatistics> runNextIteration() override {
        const size_t V = views_.getViews().size();
        std::vector<double> st(2 * V);
        check(mvd_run_iterations(views_.context(), 1, st.data()));
        ++it_;
        std::vector<IterationStatistics> out(V);
        for (size_t v = 0; v < V; ++v) { out[v].sumChange = st[2 * v]; out[v].maxChange = st[2 * v + 1]; }
        return out;
    }
};
class MultiViewDeconvolutionMul : public MultiViewDeconvolution {   // one update per iteration from all views
  public:
    using MultiViewDeconvolution::MultiViewDeconvolution;
    std::vector<IterationStatistics> runNextIteration() override {
        double st[2];
        check(mvd_run_iteration_mul(views_.context(), st));
        ++it_;
        IterationStatistics s;
        s.sumChange = st[0]; s.maxChange = st[1];
        return {s};
    }
};

// ---- operator level: one halo'd block per call (ComputeBlockSeqThread.runIteration) -------------------------------------------------------
class ComputeBlockSeqThreadB200 {
  public:
    ComputeBlockSeqThreadB200(float minValue, float lambda, int id, const Dims& blockSize, int device)
        : minValue_(minValue), lambda_(lambda), id_(id), blockSize_(blockSize), device_(device), psiBlockTmp_((size_t)numElements(blockSize)) {}
    int getId() const { return id_; }
    const Dims& getBlockSize() const { return blockSize_; }
    float getMinValue() const { return minValue_; }
    std::vector<float>& getPsiBlockTmp() { return psiBlockTmp_; }  // the caller copies the (mirror-extended) psi block in, and the result out
    IterationStatistics runIteration(const float* imgBlock, const float* weightBlock, float maxIntensityView, const DeconViewPSF& psf) {
        double st[2];
        check(mvd_block_iteration(device_, psiBlockTmp_.data(), imgBlock, weightBlock, blockSize_.data(), psf.getKernel1().data(),
                                  psf.getKernel1Dims().data(), psf.getKernel2().data(), psf.getKernel2Dims().data(), lambda_, minValue_,
                                  maxIntensityView, st));
        IterationStatistics s;
        s.sumChange = st[0]; s.maxChange = st[1];
        return s;
    }

  private:
    float minValue_, lambda_;
    int id_;
    Dims blockSize_;
    int device_;
    std::vector<float> psiBlockTmp_;
};
class ComputeBlockSeqThreadB200Factory {                            // ComputeBlockThreadFactory: create(id), numParallelBlocks()
  public:
    ComputeBlockSeqThreadB200Factory(float minValue, float lambda, const Dims& blockSize, std::vector<int> devices)
        : minValue_(minValue), lambda_(lambda), blockSize_(blockSize), devices_(std::move(devices)) {}
    ComputeBlockSeqThreadB200 create(int id) const { return ComputeBlockSeqThreadB200(minValue_, lambda_, id, blockSize_, devices_.at((size_t)id)); }
    int numParallelBlocks() const { return (int)devices_.size(); }

  private:
    float minValue_, lambda_;
    Dims blockSize_;
    std::vector<int> devices_;
};

// ---- block geometry of the reference (host logic) and the block-wise driver over the operator ---------------------------------------------
// Block (M/process/cuda/Block.java:91-133,158-239), BlockGeneratorFixedSizePrecise.divideIntoBlocks (.../BlockGeneratorFixedSizePrecise.java:59-131),
// BlockSorter.sortBlocksBySmallestFootprint (.../BlockSorter.java:55-143)
struct Block {
    Dims blockSize, offset, effectiveSize, effectiveOffset, effectiveLocalOffset;
    int min(int d) const { return offset[(size_t)d]; }
    // Block.copyBlock: cut [offset, offset + blockSize) out of `source`, extended by mirroring (extendMirrorSingle) or by zeros
    void copyBlock(const float* source, const Dims& n, bool mirror, float* block) const {
        auto map = [&](int c, int len, bool& out) {
            out = c < 0 || c >= len;
            if (!out || !mirror) return c;
            if (len == 1) return 0;
            const int period = 2 * (len - 1);
            int m = c % period;
            if (m < 0) m += period;
            return m < len ? m : period - m;
        };
        for (int z = 0; z < blockSize[2]; ++z)
            for (int y = 0; y < blockSize[1]; ++y)
                for (int x = 0; x < blockSize[0]; ++x) {
                    bool ox, oy, oz;
                    const int sx = map(offset[0] + x, n[0], ox), sy = map(offset[1] + y, n[1], oy), sz = map(offset[2] + z, n[2], oz);
                    const bool outside = ox || oy || oz;
                    block[((size_t)z * blockSize[1] + y) * blockSize[0] + x] =
                        (outside && !mirror) ? 0.f : source[((size_t)sz * n[1] + sy) * n[0] + sx];
                }
    }
    // Block.pasteBlock: the effective region only
    void pasteBlock(float* target, const Dims& n, const float* block) const {
        for (int z = 0; z < effectiveSize[2]; ++z)
            for (int y = 0; y < effectiveSize[1]; ++y)
                for (int x = 0; x < effectiveSize[0]; ++x)
                    target[((size_t)(effectiveOffset[2] + z) * n[1] + effectiveOffset[1] + y) * n[0] + effectiveOffset[0] + x] =
                        block[((size_t)(effectiveLocalOffset[2] + z) * blockSize[1] + effectiveLocalOffset[1] + y) * blockSize[0] + effectiveLocalOffset[0] + x];
    }
};

// empty result: the block is smaller than the kernel (the reference returns null)
inline std::vector<Block> divideIntoBlocks(const Dims& imgSize, const Dims& blockSize, const Dims& kernelSize) {
    Dims eff, loc, nb;
    for (size_t d = 0; d < 3; ++d) {
        eff[d] = blockSize[d] - kernelSize[d] + 1;
        if (eff[d] <= 0) return {};
        loc[d] = kernelSize[d] / 2;
        nb[d] = (imgSize[d] + eff[d] - 1) / eff[d];
    }
    std::vector<Block> blocks;
    for (int bz = 0; bz < nb[2]; ++bz)
        for (int by = 0; by < nb[1]; ++by)
            for (int bx = 0; bx < nb[0]; ++bx) {
                const Dims cur{bx, by, bz};
                Block b;
                b.blockSize = blockSize; b.effectiveLocalOffset = loc;
                for (size_t d = 0; d < 3; ++d) {
                    b.effectiveOffset[d] = cur[d] * eff[d];
                    b.effectiveSize[d] = std::min(eff[d], imgSize[d] - b.effectiveOffset[d]);
                    b.offset[d] = b.effectiveOffset[d] - loc[d];
                }
                blocks.push_back(b);
            }
    return blocks;
}

inline std::vector<std::vector<Block>> sortBlocksBySmallestFootprint(const std::vector<Block>& blocks, const Dims& psiDims, int minRequiredBlocks = 1) {
    const Dims eff = blocks.at(0).effectiveSize;
    Dims nb;
    for (size_t d = 0; d < 3; ++d) nb[d] = (psiDims[d] + eff[d] - 1) / eff[d];
    std::vector<std::pair<long long, int>> sizes;              // (blocks per layer, dimension); a later dimension overwrites an equal size
    for (int d = 0; d < 3; ++d) {
        long long sz = 1;
        for (int e = 0; e < 3; ++e) if (e != d) sz *= nb[(size_t)e];
        sizes.emplace_back(sz, d);
    }
    auto dim_of = [&](long long sz) { int dim = -1; for (const auto& p : sizes) if (p.first == sz) dim = p.second; return dim; };
    std::vector<long long> sorted{sizes[0].first, sizes[1].first, sizes[2].first};
    std::sort(sorted.begin(), sorted.end());
    int minDim = -1;
    for (int i = 0; i < 3; ++i)
        if (minDim == -1 && (sorted[(size_t)i] >= minRequiredBlocks || i == 2)) minDim = dim_of(sorted[(size_t)i]);
    std::vector<std::vector<Block>> layers;
    size_t total = 0;
    for (int i = 0; i < nb[(size_t)minDim]; ++i) {
        const int off = blocks[0].offset[(size_t)minDim] + i * eff[(size_t)minDim];
        std::vector<Block> layer;
        for (const Block& b : blocks) if (b.min(minDim) == off) layer.push_back(b);
        total += layer.size();
        layers.push_back(std::move(layer));
    }
    if (total != blocks.size()) return {blocks};
    return layers;
}

// MultiViewDeconvolutionSeq.runNextIteration through the block operator (MultiViewDeconvolutionSeq.java:69-176): per view, batches of
// non-interfering blocks, delayed paste-back of the effective regions.  psi is a caller-owned host volume, updated in place.
inline std::vector<IterationStatistics> runNextIterationBlocked(float* psi, const Dims& dims, DeconViews& views, const std::vector<float>& maxIntensities,
                                                                const ComputeBlockSeqThreadB200Factory& factory, const Dims& blockSize) {
    ComputeBlockSeqThreadB200 worker = factory.create(0);
    std::vector<IterationStatistics> out;
    const size_t nblk = (size_t)numElements(blockSize);
    std::vector<float> imgBlock(nblk), weightBlock(nblk);
    for (size_t v = 0; v < views.getViews().size(); ++v) {
        DeconView& view = views.getViews()[v];
        const Dims k1d = view.getPSF().getKernel1Dims();
        const Dims ksz{2 * k1d[0] - 1, 2 * k1d[1] - 1, 2 * k1d[2] - 1};         // DeconView.java:155-157
        const std::vector<Block> blocks = divideIntoBlocks(dims, blockSize, ksz);
        if (blocks.empty()) throw Error("block smaller than the kernel");
        IterationStatistics st;
        std::vector<std::pair<Block, std::vector<float>>> prev, cur;
        for (const std::vector<Block>& batch : sortBlocksBySmallestFootprint(blocks, dims)) {
            cur.clear();
            for (const Block& blk : batch) {
                blk.copyBlock(psi, dims, true, worker.getPsiBlockTmp().data());
                blk.copyBlock(view.getImage().data, dims, false, imgBlock.data());
                blk.copyBlock(view.getWeight().data, dims, false, weightBlock.data());
                const IterationStatistics s = worker.runIteration(imgBlock.data(), weightBlock.data(), maxIntensities.at(v), view.getPSF());
                st.sumChange += s.sumChange;
                st.maxChange = std::max(st.maxChange, s.maxChange);
                if (blocks.size() == 1) blk.pasteBlock(psi, dims, worker.getPsiBlockTmp().data());
                else cur.emplace_back(blk, worker.getPsiBlockTmp());
            }
            for (const auto& p : prev) p.first.pasteBlock(psi, dims, p.second.data());
            prev.swap(cur);
        }
        for (const auto& p : prev) p.first.pasteBlock(psi, dims, p.second.data());
        out.push_back(st);
    }
    return out;
}

}  // namespace mvrecon
#endif  // MVDECON_HPP
