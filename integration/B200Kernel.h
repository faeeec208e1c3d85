/* B200Kernel.h — the engine host class a Sol-R maintainer adds as solr/engines/b200/B200Kernel.h.
 *
 * A sibling of solr::CudaKernel (solr/engines/cuda/CudaKernel.h:28-72): same GPUKernel virtuals, same dirty-flag
 * protocol, but the device work goes through the C ABI of libsolr_b200 (include/solr_b200.h) instead of
 * CudaRayTracer.h.  It compiles against the UNMODIFIED reference headers; tests/test_integration.py builds it
 * together with the reference's own host sources and drives it with the reference's own setters.
 * INTEGRATION.md lists the three one-line edits that make SOLR_ENGINE=B200 select it.
 */
#pragma once

#include <DLL_API.h>
#include <engines/GPUKernel.h>

namespace solr
{
class SOLR_API B200Kernel : public GPUKernel
{
public:
    B200Kernel();
    ~B200Kernel();

    void initBuffers() override;
    void cleanup() override;

    void setPlatformId(const int) override {}
    void setDeviceId(const int device) override;
    void setKernelFilename(const std::string &) override {}
    void queryDevice() override;
    void recompileKernels() override {}

    void render_begin(const float timer) override;
    void render_end() override;
    std::string getGPUDescription() override { return m_gpuDescription; }

    /* frame-size limits beyond the reference's 1920x1080 (Consts.h:39-41) and the multi-GPU frame split.  GPUKernel's own frame,
     * id and random-number buffers are sized for MAX_BITMAP_SIZE and its getBitmap() / getPrimitiveAt() are not virtual, so a frame
     * beyond that size lives in buffers of this class: read it with getFrame() / getPrimitiveIdAt(). */
    void setLimits(int maxWidth, int maxHeight);
    /* the frame render_end read back, whatever its size: GPUKernel::getBitmap() within the reference's limit, else the large buffer */
    BitmapBuffer *getFrame() { return m_bigBitmap.empty() ? m_bitmap : m_bigBitmap.data(); }
    /* GPUKernel::getPrimitiveAt (GPUKernel.cpp:729-739) for any frame size */
    unsigned int getPrimitiveIdAt(int x, int y);
    void setPartition(int rank, int worldSize);
    /* The animation step on the device (include/solr_b200.h b200_rotate_primitives / b200_translate_primitives): what
     * GPUKernel::rotatePrimitives / translatePrimitives + compactBoxes(false) + the re-upload do per frame in MoleculeScene.cpp:75-81,
     * applied to the arrays where they live — same arithmetic, same arrays, no host work.  GPUKernel's own methods are not virtual
     * (GPUKernel.h:119-128,304), so a scene calls these instead while it animates, and syncFromDevice() before it reads or edits the
     * container again (getPrimitive, setPrimitive, compactBoxes(true) ...).  Both fall back to the host-side step — and return
     * false — when the scene is not on the device yet. */
    bool rotatePrimitivesOnDevice(const vec3f &rotationCenter, const vec4f &angles);
    bool translatePrimitivesOnDevice(const vec3f &translation);
    void syncFromDevice();
    /* fixes what GPUKernel::render_begin draws from rand() (GPUKernel.cpp:2719-2727): deterministic frames */
    void setRandoms(const float *randoms, size_t count, int timestamp);

private:
    void initializeDevice();
    void releaseDevice();
    bool m_deviceInitialized;
    bool m_fixedRandoms;
    bool m_hostStale; /* the last animation steps were applied on the device only */
    int m_fixedTimestamp;
    int m_maxWidth, m_maxHeight;
    std::vector<unsigned char> m_bigBitmap; /* used instead of m_bitmap when the limits exceed the reference's */
    std::vector<PrimitiveXYIdBuffer> m_bigIds;
    std::vector<float> m_bigRandoms; /* the random table at the size of the limits */
    void *m_pinned[2];               /* frame / id buffers registered with the engine, unregistered before they are freed */
    void unpin();
};
}
