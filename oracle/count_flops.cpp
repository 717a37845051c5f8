// count_flops.cpp -- instrumented FP64 operation count of the oracle's optics path.
// Compiles oracle_optics.c as C++ with `double` replaced by a counting scalar, runs
// orc_rubin_optics on a few thousand photons and prints operations per photon by class.
// This is the "algorithmic FLOPs of the reference arithmetic" used by bench.py's roofline
// (SURVEY.md section 8d asks for an instrumented count instead of the hand estimate).
// TEST INFRASTRUCTURE (oracle/); build: g++ -O1 -o _build/count_flops count_flops.cpp -lm
#include <cmath>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static unsigned long long g_add = 0, g_mul = 0, g_div = 0, g_sqrt = 0, g_trans = 0, g_cmp = 0;

struct Counted {
    double v;
    Counted() : v(0.0) {}
    Counted(double x) : v(x) {}
    Counted(int x) : v(x) {}
    Counted(long x) : v((double)x) {}
    Counted(long long x) : v((double)x) {}
    Counted(unsigned long x) : v((double)x) {}
    explicit operator double() const { return v; }
    explicit operator int() const { return (int)v; }
    explicit operator bool() const { return v != 0.0; }
    Counted operator-() const { return Counted(-v); }
    Counted& operator+=(Counted o) { g_add++; v += o.v; return *this; }
    Counted& operator-=(Counted o) { g_add++; v -= o.v; return *this; }
    Counted& operator*=(Counted o) { g_mul++; v *= o.v; return *this; }
    Counted& operator/=(Counted o) { g_div++; v /= o.v; return *this; }
};
#define BINOP(op, ctr)                                                                  \
    static inline Counted operator op(Counted a, Counted b) { ctr++; return Counted(a.v op b.v); } \
    static inline Counted operator op(Counted a, double b) { ctr++; return Counted(a.v op b); }    \
    static inline Counted operator op(double a, Counted b) { ctr++; return Counted(a op b.v); }    \
    static inline Counted operator op(Counted a, int b) { ctr++; return Counted(a.v op b); }       \
    static inline Counted operator op(int a, Counted b) { ctr++; return Counted(a op b.v); }
BINOP(+, g_add) BINOP(-, g_add) BINOP(*, g_mul) BINOP(/, g_div)
#define CMPOP(op)                                                                   \
    static inline bool operator op(Counted a, Counted b) { g_cmp++; return a.v op b.v; } \
    static inline bool operator op(Counted a, double b) { g_cmp++; return a.v op b; }    \
    static inline bool operator op(double a, Counted b) { g_cmp++; return a op b.v; }    \
    static inline bool operator op(Counted a, int b) { g_cmp++; return a.v op b; }
CMPOP(<) CMPOP(>) CMPOP(<=) CMPOP(>=) CMPOP(==) CMPOP(!=)
static inline Counted c_sqrt(Counted a) { g_sqrt++; return Counted(std::sqrt(a.v)); }
static inline Counted c_fabs(Counted a) { return Counted(std::fabs(a.v)); }
#define TRANS1(name) static inline Counted c_##name(Counted a) { g_trans++; return Counted(std::name(a.v)); }
TRANS1(sin) TRANS1(cos) TRANS1(tan) TRANS1(asin) TRANS1(atan) TRANS1(floor)
static inline Counted c_atan2(Counted a, Counted b) { g_trans++; return Counted(std::atan2(a.v, b.v)); }
static inline Counted c_pow(Counted a, int b) { for (int i = 1; i < b; ++i) g_mul++; return Counted(std::pow(a.v, b)); }
static inline Counted c_pow(Counted a, Counted b) { g_trans++; return Counted(std::pow(a.v, b.v)); }
static inline bool c_isnan(Counted a) { return std::isnan(a.v); }

// the real struct layouts (plain doubles) for building inputs
#include "../include/imsim_b200.h"
namespace plain {
typedef B2Telescope Telescope;
}

// now re-read the header and the oracle with double -> Counted inside a namespace
#undef IMSIM_B200_H
#define double Counted
#define sqrt c_sqrt
#define fabs c_fabs
#define sin c_sin
#define cos c_cos
#define tan c_tan
#define asin c_asin
#define atan c_atan
#define atan2 c_atan2
#define floor c_floor
#define pow c_pow
#define isnan c_isnan
#define NAN Counted(__builtin_nan(""))
#define INFINITY Counted(__builtin_inf())
namespace counted {
#define B2_STRUCTS_ONLY
#include "../include/imsim_b200.h"
#include "oracle_optics.c"
}  // namespace counted
#undef double
#undef sqrt
#undef fabs
#undef sin
#undef cos
#undef tan
#undef asin
#undef atan
#undef atan2
#undef floor
#undef pow
#undef isnan

int main(int argc, char** argv) {
    // inputs: raw bytes of the PODs + photon arrays written by tests/helpers (see count_flops.py)
    if (argc < 2) { fprintf(stderr, "usage: count_flops <blob>\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    static_assert(sizeof(counted::B2Telescope) == sizeof(B2Telescope), "layout");
    counted::B2Telescope tel; counted::B2TanSip img, field; counted::B2Detector det; counted::B2Diffraction dif;
    counted::B2OpticsOptions opt;
    int64_t n;
    if (fread(&tel, sizeof(tel), 1, f) != 1 || fread(&img, sizeof(img), 1, f) != 1 || fread(&field, sizeof(field), 1, f) != 1 ||
        fread(&det, sizeof(det), 1, f) != 1 || fread(&dif, sizeof(dif), 1, f) != 1 || fread(&opt, sizeof(opt), 1, f) != 1 ||
        fread(&n, sizeof(n), 1, f) != 1) return 3;
    std::vector<Counted> a[9];
    for (auto& v : a) { v.resize(n); if (fread(v.data(), sizeof(Counted), n, f) != (size_t)n) return 3; }
    fclose(f);
    std::vector<Counted> dxdz(n), dydz(n);
    counted::B2OpticsStats st;
    g_add = g_mul = g_div = g_sqrt = g_trans = g_cmp = 0;
    counted::orc_rubin_optics(&tel, nullptr, nullptr, &img, &field, &det, &dif, &opt, n, a[0].data(), a[1].data(),
                              dxdz.data(), dydz.data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(), a[6].data(),
                              a[7].data(), nullptr, &st);
    double N = (double)n;
    printf("{\"photons\": %lld, \"add\": %.1f, \"mul\": %.1f, \"div\": %.1f, \"sqrt\": %.1f, \"transcendental\": %.1f, "
           "\"compare\": %.1f, \"flop_add_mul_div_sqrt\": %.1f}\n", (long long)n, g_add / N, g_mul / N, g_div / N, g_sqrt / N,
           g_trans / N, g_cmp / N, (g_add + g_mul + g_div + g_sqrt) / N);
    return 0;
}
