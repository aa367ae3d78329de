/* CPU oracle (TEST INFRASTRUCTURE ONLY): ECE / reliability binning in plain C.
 *
 * Restates the per-bin masked passes of the reference:
 *   src/uncertainty.py:113-126  (ECE: `conf >= lower & conf < upper`, last bin `<=`)
 *   src/uncertainty.py:231-241  (reliability diagram, float64 edges)
 * One full pass over the samples per bin, exactly like the reference's
 * boolean-mask loop; comparisons are done in double (fp32 confidences and
 * fp32 edges promote exactly), so one routine serves both edge flavours.
 *
 * Parity pin: tests/test_oracle_golden.py checks this against the numpy
 * oracle and against tests/golden/ece_*.npz made from the real reference.
 */
#include <stdint.h>
#include <omp.h>

int ece_oracle_bin(const float *conf, const int64_t *pred, const int64_t *label,
                   int64_t n, const double *edges, int num_bins,
                   int64_t *count, int64_t *correct, double *conf_sum,
                   int threads)
{
    if (num_bins <= 0 || n < 0) return 1;
    if (threads < 1) threads = 1;
    for (int b = 0; b < num_bins; ++b) {
        const double lower = edges[b], upper = edges[b + 1];
        const int last = (b == num_bins - 1);
        int64_t cnt = 0, cor = 0;
        double sum = 0.0;
#pragma omp parallel for num_threads(threads) reduction(+ : cnt, cor, sum) schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const double c = (double)conf[i];
            const int in_bin = last ? (c >= lower && c <= upper)
                                    : (c >= lower && c < upper);
            if (in_bin) {
                cnt += 1;
                cor += (pred[i] == label[i]);
                sum += c;
            }
        }
        count[b] = cnt;
        correct[b] = cor;
        conf_sum[b] = sum;
    }
    return 0;
}
