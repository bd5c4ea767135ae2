#include <cstdio>
#include <cstdlib>
#include "ssb_jet.cuh"
// usage: jet_test kind x y z t  -> 20 Taylor coefficients (order 3) and the order-2 / order-1 results for cross-checking
int main(int argc, char** argv) {
    int kind = atoi(argv[1]);
    double x = atof(argv[2]), y = atof(argv[3]), z = atof(argv[4]), t = atof(argv[5]);
    double pb[8] = {4.498502151469554e-12 * 1e10, 3.5, 0.5, 0.6, 0.04, 0, 0, 0};
    double pd[8] = {0.01, 0.22, 8.0, 3.4, 0.4, 0.05, 0, 0};
    double J3[20], J2[10], J1[4];
    if (kind == 0) { ssb::bar_jet<3>(pb, pb[0], x, y, z, t, J3); ssb::bar_jet<2>(pb, pb[0], x, y, z, t, J2); ssb::bar_jet<1>(pb, pb[0], x, y, z, t, J1); }
    else { ssb::dehnen_bar_jet<3>(pd, pd[0], x, y, z, t, J3); ssb::dehnen_bar_jet<2>(pd, pd[0], x, y, z, t, J2); ssb::dehnen_bar_jet<1>(pd, pd[0], x, y, z, t, J1); }
    for (int i = 0; i < 20; ++i) printf("%.17g ", J3[i]);
    for (int i = 0; i < 10; ++i) printf("%.17g ", J2[i]);
    for (int i = 0; i < 4; ++i) printf("%.17g ", J1[i]);
    printf("\n");
}
