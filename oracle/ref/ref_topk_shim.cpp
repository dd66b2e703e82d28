// TEST INFRASTRUCTURE ONLY (oracle). Thin extern "C" door onto the
// reference's own header lvmhelpers/binary_topk.h, included from
// /root/reference where it lies (never copied; see oracle/ref/Makefile).
// Mirrors the loop of batched_topk (lvmhelpers/pbinary_topk.pyx:23-34).
#include "binary_topk.h"

extern "C" int smlvm_ref_topk(const float* scores, float* configs,
                              float* out_scores, long rows, int n, int k)
{
    for (long r = 0; r < rows; ++r) {
        std::vector<scored_cfg> out =
          topk(const_cast<float*>(scores + r * (long)n), n, k);
        for (int j = 0; j < k; ++j) {
            out[j].populate_out(configs + (r * k + j) * (long)n, n);
            if (out_scores) out_scores[r * k + j] = out[j].score;
        }
    }
    return 0;
}
