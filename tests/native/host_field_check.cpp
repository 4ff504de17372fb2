// Host-side check of csrc/field.cuh: runs the *same* carry-chain sequence the GPU runs (C
// fallback of mont_chains.cuh) on vectors read from stdin, prints results for the Python
// test (tests/test_host_field.py) to compare with the big-int oracle.
//   input : <field> <op> <count> then count lines of hex operands (a b)
//   output: one hex result per line
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>
#include "../../mpc-jellyfish_b200/csrc/field.cuh"
using namespace jf;

template <class F> static Fp<F> parse(const std::string &h) {
    Fp<F> r = Fp<F>::zero();
    int n = (int)h.size();
    for (int i = 0; i < n; i++) {
        char c = h[n - 1 - i];
        uint32_t d = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
        if (i / 8 < F::N) r.v[i / 8] |= d << (4 * (i % 8));
    }
    return r;
}
template <class F> static void show(const Fp<F> &a) {
    for (int i = F::N - 1; i >= 0; i--) printf("%08x", a.v[i]);
    printf("\n");
}
template <class F> static int run(const std::string &op, int count) {
    if (op == "consts") {
        show(Fp<F>::one()); show(Fp<F>::r_squared());
        Fp<F> p; Limbs<F>::p(p.v); show(p); printf("%08x\n", F::INV);
        return 0;
    }
    for (int i = 0; i < count; i++) {
        std::string sa, sb;
        std::cin >> sa >> sb;
        Fp<F> a = parse<F>(sa), b = parse<F>(sb), r;
        if (op == "mul") r = Fp<F>::mul(a, b);
        else if (op == "madd") r = Fp<F>::mul_add(a, b, Fp<F>::add(a, b), Fp<F>::sub(a, b));  // a b + (a + b)(a - b), one reduction
        else if (op == "sqr") {  // the dedicated squaring where the field has the headroom for it, else what sqr() does
            if constexpr (Fp<F>::SQR_OK) r = Fp<F>::sqr_dedicated(a);
            else r = Fp<F>::sqr(a);
        }
        else if (op == "add") r = Fp<F>::add(a, b);
        else if (op == "sub") r = Fp<F>::sub(a, b);
        else if (op == "neg") r = Fp<F>::neg(a);
        else if (op == "inv") r = Fp<F>::inv(a);
        else if (op == "to_mont") r = Fp<F>::to_mont(a);
        else if (op == "from_mont") r = Fp<F>::from_mont(a);
        else return 2;
        show(r);
    }
    return 0;
}
int main() {
    std::string field, op;
    int count;
    std::cin >> field >> op >> count;
    if (field == "bn254_fr") return run<Bn254Fr>(op, count);
    if (field == "bn254_fq") return run<Bn254Fq>(op, count);
    if (field == "bls12_381_fr") return run<Bls12381Fr>(op, count);
    if (field == "bls12_381_fq") return run<Bls12381Fq>(op, count);
    return 1;
}
