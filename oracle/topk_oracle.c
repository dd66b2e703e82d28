/* TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
 *
 * CPU restatement of the reference's k-best bit-vector search
 * (lvmhelpers/binary_topk.h:57-129, driven per row by
 * lvmhelpers/pbinary_topk.pyx:23-34). Pinned against the reference itself:
 * oracle/ref/Makefile compiles binary_topk.h + pbinary_topk.pyx from
 * /root/reference into oracle/_ref/ and tests/test_oracle_ref.py compares.
 *
 * Algorithm (zeroth-order k-best Viterbi over independent bits):
 *   list = [(x0,{0}), (0,{})] ordered by x0 >= 0           (binary_topk.h:104-116)
 *   for bit i = 1..n-1: merge list A (bit i off) with A + x_i (bit i on),
 *     taking the "on" item only when its fp32 score is STRICTLY greater
 *     (binary_topk.h:72), keep at most k items                (:57-98)
 *   scores accumulate in fp32, in bit order                    (:27-29)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SMLVM_TOPK_MAX_BITS 512            /* binary_topk.h:13 */
#define WORDS (SMLVM_TOPK_MAX_BITS / 32)

typedef struct {
    float score;
    uint32_t bits[WORDS];
} item_t;

/* One row. out_scores[k], out_bits[k][words] (bit i of a config = bit i%32 of
 * word i/32). Returns the number of items produced (= k when 2^n >= k). */
static int topk_row(const float* x, int n, int k, item_t* cur, item_t* nxt)
{
    int words = (n + 31) / 32;
    item_t c0, c1;
    memset(&c0, 0, sizeof c0);
    c1 = c0;
    c1.score = c0.score + x[0];
    c1.bits[0] |= 1u;
    if (x[0] >= 0) { cur[0] = c1; cur[1] = c0; }
    else           { cur[0] = c0; cur[1] = c1; }
    int len = 2;
    (void)words;
    for (int i = 1; i < n; ++i) {
        float v = x[i];
        int a = 0, b = 0, m = 0;
        while (m < k && a < len && b < len) {
            float bs = cur[b].score + v;       /* fp32 add, as `update` */
            if (bs > cur[a].score) {
                nxt[m] = cur[b];
                nxt[m].score = bs;
                nxt[m].bits[i >> 5] |= 1u << (i & 31);
                ++b;
            } else {
                nxt[m] = cur[a];
                ++a;
            }
            ++m;
        }
        while (m < k && a < len) { nxt[m++] = cur[a++]; }
        while (m < k && b < len) {
            nxt[m] = cur[b];
            nxt[m].score = cur[b].score + v;
            nxt[m].bits[i >> 5] |= 1u << (i & 31);
            ++b; ++m;
        }
        item_t* tmp = cur; cur = nxt; nxt = tmp;
        len = m;
    }
    /* result is in `cur`; caller alternates buffers the same way */
    return len;
}

/* Batched: scores[rows][n] fp32 -> configs[rows][k][n] fp32 in {0,1}
 * (exactly the in-place contract of batched_topk, pbinary_topk.pyx:23-34),
 * optional out_scores[rows][k] (the DP's own fp32 scores).
 * Returns 0, -1 on bad arguments (k < 2, n < 1, n > 512, 2^n < k). */
int smlvm_oracle_topk(const float* scores, float* configs, float* out_scores,
                      long rows, int n, int k)
{
    if (k < 2 || n < 1 || n > SMLVM_TOPK_MAX_BITS) return -1;
    if (n < 31 && (1L << n) < k) return -1;
    item_t* buf0 = (item_t*)malloc(sizeof(item_t) * (size_t)k);
    item_t* buf1 = (item_t*)malloc(sizeof(item_t) * (size_t)k);
    if (!buf0 || !buf1) { free(buf0); free(buf1); return -2; }
    for (long r = 0; r < rows; ++r) {
        const float* x = scores + r * (long)n;
        topk_row(x, n, k, buf0, buf1);
        /* n-1 swaps happened inside topk_row */
        item_t* res = ((n - 1) & 1) ? buf1 : buf0;
        for (int j = 0; j < k; ++j) {
            float* dst = configs + ((r * k) + j) * (long)n;
            for (int i = 0; i < n; ++i)
                dst[i] = (float)((res[j].bits[i >> 5] >> (i & 31)) & 1u);
            if (out_scores) out_scores[r * k + j] = res[j].score;
        }
    }
    free(buf0);
    free(buf1);
    return 0;
}
