// scatter_emul.cpp -- TEST HARNESS (CPU): the ordered density scatter of F2D_SEM_CPU executed on the host with the
// very per-cell functions the CUDA kernels call (fluid-2d_b200/csrc/f2d_scatter_core.h): pass 1 = k_scatter_keys,
// pass 2 = k_scatter_ordered, one "thread" per cell in an arbitrary order (the result must not depend on it).
// tests/test_scatter_core_cpu.py compares the interior with the oracle's sequential scatter bit for bit.
// This is not a CPU fallback of the product: libf2d.so never contains or calls it.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../fluid-2d_b200/csrc/f2d_scatter_core.h"

using namespace f2d::sc;

extern "C" void scatter_emul(const float* src, const float* u, const float* v, float* out, int rows, int cols, int pitch,
                             float dt0, int reverse_order) {
    std::vector<unsigned> keys((size_t)rows * pitch, 12345u);
    unsigned disp = 0;
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) {
            const size_t o = (size_t)i * pitch + j;
            unsigned key = kNoKey;
            if (i >= 1 && i <= rows - 2 && j >= 1 && j <= cols - 2) {
                float a = fmul(dt0, u[o]), b = fmul(dt0, v[o]);
                if (a < 0) a = -a;
                if (b < 0) b = -b;
                unsigned ua, ub;
                memcpy(&ua, &a, 4);
                memcpy(&ub, &b, 4);
                disp = std::max(disp, std::max(ua, ub));
                key = source_key(rows, cols, pitch, i, j, u[o], v[o], dt0);
            }
            keys[o] = key;
        }
    const int R = reach(disp, std::max(rows, cols));
    const unsigned P = (unsigned)pitch;
    const int nt = (rows - 2) * (cols - 2);
    for (int t = 0; t < nt; ++t) {
        const int tt = reverse_order ? nt - 1 - t : t;  // thread order is irrelevant: every target is independent
        const int ti = 1 + tt / (cols - 2), tj = 1 + tt % (cols - 2);
        const int ilo = std::max(1, ti - R), ihi = std::min(rows - 2, ti + R);
        const int jlo = std::max(1, tj - R), jhi = std::min(cols - 2, tj + R);
        const unsigned T = (unsigned)ti * P + (unsigned)tj;
        float acc = 0.f;
        for (int i = ilo; i <= ihi; ++i)
            for (int j = jlo; j <= jhi; ++j) {
                const size_t o = (size_t)i * pitch + j;
                const unsigned d = T - keys[o];
                if (!is_hit(d, P)) continue;
                acc = fadd(acc, share(rows, cols, i, j, u[o], v[o], dt0, d, P, src[o]));
            }
        out[(size_t)ti * pitch + tj] = acc;
    }
}
