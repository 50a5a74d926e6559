// Stage-wise arithmetic on N independent chains, each stage ONE volatile asm block.   (generated: tools/gen_ilp_asm.py)
//
// Why: nvcc's front end orders independent dependent-chains depth first (to save registers), and ptxas keeps - or, under
// register pressure, restores - that order: the chains of N pair evaluations then issue one after the other, each
// instruction waiting for its predecessor (~12 cycles per FP64 operation, two issue cycles of work).  Volatile asm
// statements keep their relative order, so a block per stage pins "stage s of all chains before stage s+1 of any".
// Every operation is written with an explicit rounding mode: nothing here can be contracted or re-associated.
#pragma once

namespace sol {
namespace ilp {

template <int N> struct V;

template <> struct V<2> {
	static __device__ __forceinline__ void sub_sv(double (&d)[2], const double s, const double (&v)[2])
	{ asm volatile("sub.rn.f64 %0, %2, %3;\n\tsub.rn.f64 %1, %2, %4;" : "=&d"(d[0]), "=&d"(d[1]) : "d"(s), "d"(v[0]), "d"(v[1])); }
	static __device__ __forceinline__ void sub_vs(double (&d)[2], const double (&v)[2], const double s)
	{ asm volatile("sub.rn.f64 %0, %3, %2;\n\tsub.rn.f64 %1, %4, %2;" : "=&d"(d[0]), "=&d"(d[1]) : "d"(s), "d"(v[0]), "d"(v[1])); }
	static __device__ __forceinline__ void mul_vv(double (&d)[2], const double (&a)[2], const double (&b)[2])
	{ asm volatile("mul.rn.f64 %0, %2, %4;\n\tmul.rn.f64 %1, %3, %5;" : "=&d"(d[0]), "=&d"(d[1]) : "d"(a[0]), "d"(a[1]), "d"(b[0]), "d"(b[1])); }
	static __device__ __forceinline__ void mul_sv(double (&d)[2], const double s, const double (&v)[2])
	{ asm volatile("mul.rn.f64 %0, %2, %3;\n\tmul.rn.f64 %1, %2, %4;" : "=&d"(d[0]), "=&d"(d[1]) : "d"(s), "d"(v[0]), "d"(v[1])); }
	static __device__ __forceinline__ void fma_vvv(double (&d)[2], const double (&a)[2], const double (&b)[2], const double (&c)[2])
	{ asm volatile("fma.rn.f64 %0, %2, %4, %6;\n\tfma.rn.f64 %1, %3, %5, %7;" : "=&d"(d[0]), "=&d"(d[1]) : "d"(a[0]), "d"(a[1]), "d"(b[0]), "d"(b[1]), "d"(c[0]), "d"(c[1])); }
	static __device__ __forceinline__ void fma_sq_acc(double (&d)[2], const double (&a)[2])
	{ asm volatile("fma.rn.f64 %0, %2, %2, %0;\n\tfma.rn.f64 %1, %3, %3, %1;" : "+d"(d[0]), "+d"(d[1]) : "d"(a[0]), "d"(a[1])); }
	static __device__ __forceinline__ void fma_acc(double (&d)[2], const double (&a)[2], const double (&b)[2])
	{ asm volatile("fma.rn.f64 %0, %2, %4, %0;\n\tfma.rn.f64 %1, %3, %5, %1;" : "+d"(d[0]), "+d"(d[1]) : "d"(a[0]), "d"(a[1]), "d"(b[0]), "d"(b[1])); }
	static __device__ __forceinline__ void fma_vvs(double (&d)[2], const double (&a)[2], const double (&b)[2], const double s)
	{ asm volatile("fma.rn.f64 %0, %2, %4, %6;\n\tfma.rn.f64 %1, %3, %5, %6;" : "=&d"(d[0]), "=&d"(d[1]) : "d"(a[0]), "d"(a[1]), "d"(b[0]), "d"(b[1]), "d"(s)); }
	static __device__ __forceinline__ void fma_svs(double (&d)[2], const double s1, const double (&v)[2], const double s2)
	{ asm volatile("fma.rn.f64 %0, %2, %4, %3;\n\tfma.rn.f64 %1, %2, %5, %3;" : "=&d"(d[0]), "=&d"(d[1]) : "d"(s1), "d"(s2), "d"(v[0]), "d"(v[1])); }
	static __device__ __forceinline__ void rsqrt(double (&d)[2], const double (&a)[2])
	{ asm volatile("rsqrt.approx.ftz.f64 %0, %2;\n\trsqrt.approx.ftz.f64 %1, %3;" : "=&d"(d[0]), "=&d"(d[1]) : "d"(a[0]), "d"(a[1])); }
};

template <> struct V<4> {
	static __device__ __forceinline__ void sub_sv(double (&d)[4], const double s, const double (&v)[4])
	{ asm volatile("sub.rn.f64 %0, %4, %5;\n\tsub.rn.f64 %1, %4, %6;\n\tsub.rn.f64 %2, %4, %7;\n\tsub.rn.f64 %3, %4, %8;" : "=&d"(d[0]), "=&d"(d[1]), "=&d"(d[2]), "=&d"(d[3]) : "d"(s), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])); }
	static __device__ __forceinline__ void sub_vs(double (&d)[4], const double (&v)[4], const double s)
	{ asm volatile("sub.rn.f64 %0, %5, %4;\n\tsub.rn.f64 %1, %6, %4;\n\tsub.rn.f64 %2, %7, %4;\n\tsub.rn.f64 %3, %8, %4;" : "=&d"(d[0]), "=&d"(d[1]), "=&d"(d[2]), "=&d"(d[3]) : "d"(s), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])); }
	static __device__ __forceinline__ void mul_vv(double (&d)[4], const double (&a)[4], const double (&b)[4])
	{ asm volatile("mul.rn.f64 %0, %4, %8;\n\tmul.rn.f64 %1, %5, %9;\n\tmul.rn.f64 %2, %6, %10;\n\tmul.rn.f64 %3, %7, %11;" : "=&d"(d[0]), "=&d"(d[1]), "=&d"(d[2]), "=&d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3])); }
	static __device__ __forceinline__ void mul_sv(double (&d)[4], const double s, const double (&v)[4])
	{ asm volatile("mul.rn.f64 %0, %4, %5;\n\tmul.rn.f64 %1, %4, %6;\n\tmul.rn.f64 %2, %4, %7;\n\tmul.rn.f64 %3, %4, %8;" : "=&d"(d[0]), "=&d"(d[1]), "=&d"(d[2]), "=&d"(d[3]) : "d"(s), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])); }
	static __device__ __forceinline__ void fma_vvv(double (&d)[4], const double (&a)[4], const double (&b)[4], const double (&c)[4])
	{ asm volatile("fma.rn.f64 %0, %4, %8, %12;\n\tfma.rn.f64 %1, %5, %9, %13;\n\tfma.rn.f64 %2, %6, %10, %14;\n\tfma.rn.f64 %3, %7, %11, %15;" : "=&d"(d[0]), "=&d"(d[1]), "=&d"(d[2]), "=&d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3])); }
	static __device__ __forceinline__ void fma_sq_acc(double (&d)[4], const double (&a)[4])
	{ asm volatile("fma.rn.f64 %0, %4, %4, %0;\n\tfma.rn.f64 %1, %5, %5, %1;\n\tfma.rn.f64 %2, %6, %6, %2;\n\tfma.rn.f64 %3, %7, %7, %3;" : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3])); }
	static __device__ __forceinline__ void fma_acc(double (&d)[4], const double (&a)[4], const double (&b)[4])
	{ asm volatile("fma.rn.f64 %0, %4, %8, %0;\n\tfma.rn.f64 %1, %5, %9, %1;\n\tfma.rn.f64 %2, %6, %10, %2;\n\tfma.rn.f64 %3, %7, %11, %3;" : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3])); }
	static __device__ __forceinline__ void fma_vvs(double (&d)[4], const double (&a)[4], const double (&b)[4], const double s)
	{ asm volatile("fma.rn.f64 %0, %4, %8, %12;\n\tfma.rn.f64 %1, %5, %9, %12;\n\tfma.rn.f64 %2, %6, %10, %12;\n\tfma.rn.f64 %3, %7, %11, %12;" : "=&d"(d[0]), "=&d"(d[1]), "=&d"(d[2]), "=&d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]), "d"(s)); }
	static __device__ __forceinline__ void fma_svs(double (&d)[4], const double s1, const double (&v)[4], const double s2)
	{ asm volatile("fma.rn.f64 %0, %4, %6, %5;\n\tfma.rn.f64 %1, %4, %7, %5;\n\tfma.rn.f64 %2, %4, %8, %5;\n\tfma.rn.f64 %3, %4, %9, %5;" : "=&d"(d[0]), "=&d"(d[1]), "=&d"(d[2]), "=&d"(d[3]) : "d"(s1), "d"(s2), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])); }
	static __device__ __forceinline__ void rsqrt(double (&d)[4], const double (&a)[4])
	{ asm volatile("rsqrt.approx.ftz.f64 %0, %4;\n\trsqrt.approx.ftz.f64 %1, %5;\n\trsqrt.approx.ftz.f64 %2, %6;\n\trsqrt.approx.ftz.f64 %3, %7;" : "=&d"(d[0]), "=&d"(d[1]), "=&d"(d[2]), "=&d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3])); }
};

}  // namespace ilp
}  // namespace sol
