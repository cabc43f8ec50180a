// ORACLE (test infrastructure). C-callable shim around the REFERENCE's own polygon IoU, compiled from the source
// where it lies (/root/reference/tools/prepare_dota/polyiou.cpp -- never copied into this repository) into
// oracle/_ref/libpolyiou_ref.so by oracle/Makefile. Used only to validate oracle/polyiou_oracle.c.
#include <vector>
double iou_poly(std::vector<double> p, std::vector<double> q);  // tools/prepare_dota/polyiou.cpp:108
extern "C" double ref_iou_poly(const double* p, const double* q) {
    return iou_poly(std::vector<double>(p, p + 8), std::vector<double>(q, q + 8));
}
extern "C" void ref_iou_poly_batch(const double* p, const double* q, double* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = ref_iou_poly(p + 8 * i, q + 8 * i);
}
