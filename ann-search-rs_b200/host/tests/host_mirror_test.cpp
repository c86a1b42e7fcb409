// Compile/link check of the C++ host mirror; on a box without a GPU every compute call must fail with the Cuda variant
// (no CPU fallback); on a GPU box it runs the reference's 5-point fixture (src/cpu/exhaustive.rs:319-533).
#include <cmath>
#include <cstdio>

#include "../annb200.hpp"

int main() {
    using namespace annb200;
    const float data[15] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 0, 1, 0, 1};
    const float q[3] = {1, 0, 0};
    try {
        auto idx = build_exhaustive_index_gpu(MatRef::row_major(data, 5, 3), "cosine");
        auto res = query_exhaustive_index_gpu(MatRef::row_major(q, 1, 3), idx, 5, true);
        const float want[5] = {0.f, 1.f - 1.f / std::sqrt(2.f), 1.f - 1.f / std::sqrt(2.f), 1.f, 1.f};
        if (res.indices[0].size() != 5 || res.indices[0][0] != 0) { std::printf("FAIL ids\n"); return 1; }
        for (int i = 0; i < 5; i++)
            if (std::fabs((*res.distances)[0][i] - want[i]) > 1e-5f) { std::printf("FAIL dist %d\n", i); return 1; }
        // column-major input goes through matrix_to_flat
        const float cm[15] = {1, 0, 0, 1, 1, 0, 1, 0, 1, 0, 0, 0, 1, 0, 1};
        auto idx2 = build_exhaustive_index_gpu(MatRef::col_major(cm, 5, 3), "euclidean");
        auto r2 = query_exhaustive_index_gpu_self(idx2, 1, false);
        for (size_t i = 0; i < 5; i++) if (r2.indices[i][0] != i) { std::printf("FAIL self\n"); return 1; }
        try { build_exhaustive_index_gpu(MatRef::row_major(data, 5, 3), "manhattan"); std::printf("FAIL manhattan accepted\n"); return 1; }
        catch (const AnnSearchError& e) { if (e.variant() != ErrorVariant::DistanceNotSupported) return 1; }
        std::printf("OK gpu\n");
        return 0;
    } catch (const AnnSearchError& e) {
        if (e.variant() == ErrorVariant::Cuda) { std::printf("OK no-gpu: %s\n", e.what()); return 0; }
        std::printf("FAIL unexpected error: %s\n", e.what());
        return 1;
    }
}
