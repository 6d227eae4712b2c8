// tests/host_check/frame_pipe_host.cpp — TEST INFRASTRUCTURE: drives the product's FramePipe (csrc/frame_pipe.cu, compiled
// unchanged against the emulated runtime of tests/host_check/emu/cuda_runtime.h) the way Solver::step does: acquire a device
// slot, "launch the export kernel" on the solver stream (here: an operation that fills the slot with a pattern of the step, and
// the slot's scalars), submit; then drain and read every published frame back.
#include "frame_pipe.h"
#include <cstdlib>

using namespace vfd;

static inline float pattern(uint32_t step, uint32_t p, uint32_t k) { return (float)(step * 1000003u + p * 9u + k); }

// keepMask: bit (step % 32) set = the "device" marks step as a frame (only consulted when conditional != 0).
// Returns the number of published frames, or a negative code: -1 configure/submit error, -2 a frame's content is wrong,
// -3 frames out of order, -4 the zero-copy view differs from the copy.
extern "C" long hc_frame_pipe_bake(uint32_t n, uint32_t steps, int conditional, uint32_t keepMask, int exportMicros, int copyMicros,
                                   long* pinnedAllocs, long* pinnedBytes, uint32_t* publishedSteps) {
    g_emuCopyMicros = copyMicros;
    g_emuPinnedAllocs = 0; g_emuPinnedBytes = 0;
    long result = 0;
    {
        FramePipe pipe;
        cudaStream_t solver;
        cudaStreamCreateWithFlags(&solver, cudaStreamNonBlocking);
        if (pipe.configure(0, n) != cudaSuccess) return -1;
        for (uint32_t s = 0; s < steps; s++) {
            VfdParticleSimple* d = pipe.acquire();
            if (!d) return -1;
            float* meta = pipe.meta_slot();
            const bool keep = !conditional || ((keepMask >> (s % 32u)) & 1u);
            solver->push([=] {                               // the export kernel of step s
                std::this_thread::sleep_for(std::chrono::microseconds(exportMicros));
                float* f = reinterpret_cast<float*>(d);
                for (uint32_t p = 0; p < n; p++) for (uint32_t k = 0; k < 9; k++) f[9 * (size_t)p + k] = pattern(s, p, k);
                meta[0] = (float)s + 0.5f; meta[1] = (float)s + 0.25f; meta[2] = keep ? 1.0f : 0.0f; meta[3] = 0.0f;
            });
            if (pipe.submit(solver, -1.0f, -1.0f, true, conditional != 0) != cudaSuccess) return -1;
        }
        if (pipe.drain() != cudaSuccess) return -1;
        const size_t pub = pipe.published();
        std::vector<VfdParticleSimple> out(n ? n : 1);
        uint32_t expectStep = 0;
        for (size_t i = 0; i < pub && result == 0; i++) {
            float mv = 0.0f, dt = 0.0f;
            if (!pipe.read((uint32_t)i, out.data(), &mv, &dt)) { result = -2; break; }
            const uint32_t step = (uint32_t)(mv - 0.5f);
            while (conditional && expectStep < steps && !((keepMask >> (expectStep % 32u)) & 1u)) expectStep++;
            if (step != expectStep || dt != (float)step + 0.25f) { result = -3; break; }
            expectStep++;
            if (publishedSteps) publishedSteps[i] = step;
            const float* f = reinterpret_cast<const float*>(out.data());
            for (uint32_t p = 0; p < n && result == 0; p++) for (uint32_t k = 0; k < 9; k++) if (f[9 * (size_t)p + k] != pattern(step, p, k)) { result = -2; break; }
            const VfdParticleSimple* v = nullptr; uint32_t cnt = 0;
            if (!pipe.view((uint32_t)i, &v, &cnt, nullptr, nullptr) || cnt != n || memcmp(v, out.data(), (size_t)n * sizeof(VfdParticleSimple)) != 0) result = -4;
        }
        if (result == 0) result = (long)pub;
        // a second bake on the same pipe: the storage of the first is reused (clear() keeps the pool)
        pipe.clear();
        if (pipe.published() != 0) result = -3;
        cudaStreamDestroy(solver);
    }
    if (pinnedAllocs) *pinnedAllocs = g_emuPinnedAllocs;
    if (pinnedBytes) *pinnedBytes = g_emuPinnedBytes;
    return result;
}
