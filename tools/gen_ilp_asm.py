"""Generates solaris_b200/csrc/ilp_asm.cuh: stage-wise arithmetic on N independent chains, one volatile asm block per
stage (see the header's comment).  python tools/gen_ilp_asm.py"""
import os

HEAD = '''// Stage-wise arithmetic on N independent chains, each stage ONE volatile asm block.   (generated: tools/gen_ilp_asm.py)
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
'''


def struct(n):
    def outs(name): return ", ".join(f'"=&d"({name}[{k}])' for k in range(n))   # early clobber: written before all inputs are read
    def inouts(name): return ", ".join(f'"+d"({name}[{k}])' for k in range(n))
    def ins(name): return ", ".join(f'"d"({name}[{k}])' for k in range(n))
    def body(f): return "\\n\\t".join(f(k) for k in range(n))
    fn = "\tstatic __device__ __forceinline__ void "
    L = [f"template <> struct V<{n}> {{"]
    L.append(f'{fn}sub_sv(double (&d)[{n}], const double s, const double (&v)[{n}])\n\t{{ asm volatile("' + body(lambda k: f"sub.rn.f64 %{k}, %{n}, %{n+1+k};") + f'" : {outs("d")} : "d"(s), {ins("v")}); }}')
    L.append(f'{fn}sub_vs(double (&d)[{n}], const double (&v)[{n}], const double s)\n\t{{ asm volatile("' + body(lambda k: f"sub.rn.f64 %{k}, %{n+1+k}, %{n};") + f'" : {outs("d")} : "d"(s), {ins("v")}); }}')
    L.append(f'{fn}mul_vv(double (&d)[{n}], const double (&a)[{n}], const double (&b)[{n}])\n\t{{ asm volatile("' + body(lambda k: f"mul.rn.f64 %{k}, %{n+k}, %{2*n+k};") + f'" : {outs("d")} : {ins("a")}, {ins("b")}); }}')
    L.append(f'{fn}mul_sv(double (&d)[{n}], const double s, const double (&v)[{n}])\n\t{{ asm volatile("' + body(lambda k: f"mul.rn.f64 %{k}, %{n}, %{n+1+k};") + f'" : {outs("d")} : "d"(s), {ins("v")}); }}')
    L.append(f'{fn}fma_vvv(double (&d)[{n}], const double (&a)[{n}], const double (&b)[{n}], const double (&c)[{n}])\n\t{{ asm volatile("' + body(lambda k: f"fma.rn.f64 %{k}, %{n+k}, %{2*n+k}, %{3*n+k};") + f'" : {outs("d")} : {ins("a")}, {ins("b")}, {ins("c")}); }}')
    L.append(f'{fn}fma_sq_acc(double (&d)[{n}], const double (&a)[{n}])\n\t{{ asm volatile("' + body(lambda k: f"fma.rn.f64 %{k}, %{n+k}, %{n+k}, %{k};") + f'" : {inouts("d")} : {ins("a")}); }}')
    L.append(f'{fn}fma_acc(double (&d)[{n}], const double (&a)[{n}], const double (&b)[{n}])\n\t{{ asm volatile("' + body(lambda k: f"fma.rn.f64 %{k}, %{n+k}, %{2*n+k}, %{k};") + f'" : {inouts("d")} : {ins("a")}, {ins("b")}); }}')
    L.append(f'{fn}fma_vvs(double (&d)[{n}], const double (&a)[{n}], const double (&b)[{n}], const double s)\n\t{{ asm volatile("' + body(lambda k: f"fma.rn.f64 %{k}, %{n+k}, %{2*n+k}, %{3*n};") + f'" : {outs("d")} : {ins("a")}, {ins("b")}, "d"(s)); }}')
    L.append(f'{fn}fma_svs(double (&d)[{n}], const double s1, const double (&v)[{n}], const double s2)\n\t{{ asm volatile("' + body(lambda k: f"fma.rn.f64 %{k}, %{n}, %{n+2+k}, %{n+1};") + f'" : {outs("d")} : "d"(s1), "d"(s2), {ins("v")}); }}')
    L.append(f'{fn}rsqrt(double (&d)[{n}], const double (&a)[{n}])\n\t{{ asm volatile("' + body(lambda k: f"rsqrt.approx.ftz.f64 %{k}, %{n+k};") + f'" : {outs("d")} : {ins("a")}); }}')
    L.append("};\n")
    return "\n".join(L)


def generate() -> str:
    return HEAD + "\n" + "\n".join(struct(n) for n in (2, 4)) + "\n}  // namespace ilp\n}  // namespace sol\n"


PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "solaris_b200", "csrc", "ilp_asm.cuh")

if __name__ == "__main__":
    open(PATH, "w").write(generate())
    print(PATH)
