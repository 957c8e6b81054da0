// TEST INFRASTRUCTURE ONLY.  Thin extern "C" shim around the reference's own,
// unmodified host checker (verify.hpp), compiled from where it lies:
//   g++ -O2 -std=c++17 -I/root/reference/parallel_pivot -shared -fPIC ref_verify_shim.cpp \
//       -o _ref/libref_verify.so
// (no -fopenmp: the reference's `#pragma omp parallel for` in verifyInv races on its
// shared counters).  Exposes
//   verifyInv  (parallel_pivot/verify.hpp:50-103)  -> counts parsed from its stdout lines
//   pivotedA   (parallel_pivot/verify.hpp:106-155) -> permuted matrix + pivot vector
//   calc_cond_num (parallel_pivot/verify.hpp:245-338) -> the value main() prints
//   verifyLUwithPivoting (parallel_pivot/verify.hpp:157-242) -> counts parsed from its stdout lines
//     (A = ONE already permuted n x n matrix, as main() would pass pivotedA's output; LU = batch factors)
// so tests can pin the oracle's restatement against the real code.
#include <cstdint>
#include <sstream>
#include <string>
#include "verify.hpp"

namespace {
template <typename T>
void run_verify(const T* A, const T* Ainv, int n, int batch, long long* ok, long long* bad) {
    std::vector<T> a(A, A + (size_t)n * n * batch), x(Ainv, Ainv + (size_t)n * n * batch);
    std::ostringstream cap;
    std::streambuf* old = std::cout.rdbuf(cap.rdbuf());
    verifyInv<T>(a, x, n, batch);
    std::cout.rdbuf(old);
    long long c = -1, w = -1;
    std::istringstream in(cap.str());
    std::string line;
    while (std::getline(in, line)) {
        if (line.rfind("Correct inversions: ", 0) == 0) c = std::stoll(line.substr(20));
        if (line.rfind("Incorrect inversions: ", 0) == 0) w = std::stoll(line.substr(22));
    }
    *ok = c; *bad = w;
}
template <typename T>
void run_verify_lu(const T* PA, const T* LU, int n, int batch, long long* ok, long long* bad) {
    std::vector<T> a(PA, PA + (size_t)n * n), lu(LU, LU + (size_t)n * n * batch);
    std::ostringstream cap;
    std::streambuf* old = std::cout.rdbuf(cap.rdbuf());
    verifyLUwithPivoting<T>(a, lu, n, batch);
    std::cout.rdbuf(old);
    long long c = -1, w = -1;
    std::istringstream in(cap.str());
    std::string line;
    while (std::getline(in, line)) {
        if (line.rfind("Correct LU decompositions: ", 0) == 0) c = std::stoll(line.substr(27));
        if (line.rfind("Incorrect LU decompositions: ", 0) == 0) w = std::stoll(line.substr(29));
    }
    *ok = c; *bad = w;
}
template <typename T>
void run_pivoted(const T* A, T* PA, int32_t* piv, int n) {
    std::vector<T> a(A, A + (size_t)n * n), pa((size_t)n * n);
    std::vector<int> p(n, 0);
    pivotedA<T>(a, pa, p, n);
    for (int i = 0; i < n * n; ++i) PA[i] = pa[i];
    for (int i = 0; i < n; ++i) piv[i] = p[i];
}
}  // namespace

extern "C" {
void ref_verify_inv_f32(const float* A, const float* X, int n, int batch, long long* ok, long long* bad) { run_verify<float>(A, X, n, batch, ok, bad); }
void ref_verify_inv_f64(const double* A, const double* X, int n, int batch, long long* ok, long long* bad) { run_verify<double>(A, X, n, batch, ok, bad); }
void ref_verify_lu_piv_f32(const float* PA, const float* LU, int n, int batch, long long* ok, long long* bad) { run_verify_lu<float>(PA, LU, n, batch, ok, bad); }
void ref_verify_lu_piv_f64(const double* PA, const double* LU, int n, int batch, long long* ok, long long* bad) { run_verify_lu<double>(PA, LU, n, batch, ok, bad); }
void ref_write_to_file_f32(const float* A, const char* path, int n, int batch) { std::vector<float> a(A, A + (size_t)n * n * batch); writeToFile<float>(a, path, n, batch); }
void ref_write_to_file_f64(const double* A, const char* path, int n, int batch) { std::vector<double> a(A, A + (size_t)n * n * batch); writeToFile<double>(a, path, n, batch); }
void ref_pivotedA_f32(const float* A, float* PA, int32_t* piv, int n) { run_pivoted<float>(A, PA, piv, n); }
void ref_pivotedA_f64(const double* A, double* PA, int32_t* piv, int n) { run_pivoted<double>(A, PA, piv, n); }
double ref_calc_cond_num_f32(const float* A, int n) { std::vector<float> a(A, A + (size_t)n * n); return (double)calc_cond_num<float>(a, n); }
double ref_calc_cond_num_f64(const double* A, int n) { std::vector<double> a(A, A + (size_t)n * n); return calc_cond_num<double>(a, n); }
}
