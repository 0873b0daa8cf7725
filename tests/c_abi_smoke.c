/* A plain C consumer of include/custos_b200.h — the view a Rust/C/Go binding has of the library.
 * Builds the expression x.mul(2).add(1).sin() as cb_node[], runs it on the GPU through the C ABI and
 * checks the result against libm.  Exit codes: 0 ok, 3 no CUDA device (CB_ERR_NO_DEVICE), 1 failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "custos_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int32_t rc_ = (call);                                                    \
        if (rc_ != CB_OK) {                                                      \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, cb_last_error());      \
            return rc_ == CB_ERR_NO_DEVICE ? 3 : 1;                              \
        }                                                                        \
    } while (0)

int main(void)
{
    enum { N = 100003 };
    cb_device *dev = NULL;
    CHECK(cb_device_create(0, &dev));

    /* x.mul(2.0).add(1.0).sin(): nodes in topological order */
    cb_node prog[6] = {
        {CB_OP_X, -1, -1, 0, 0.0, 0},   {CB_OP_CONST, -1, -1, 0, 2.0, 0}, {CB_OP_MUL, 0, 1, 0, 0.0, 0},
        {CB_OP_CONST, -1, -1, 0, 1.0, 0}, {CB_OP_ADD, 2, 3, 0, 0.0, 0},   {CB_OP_SIN, 4, -1, 0, 0.0, 0}};
    const cb_node *progs[1] = {prog};
    const int32_t counts[1] = {6};
    char src[256];
    CHECK(cb_expr_to_cl_source(CB_F32, prog, 6, "x", "y", src, sizeof src));
    printf("to_cl_source: %s\n", src);

    cb_expr *f = NULL;
    CHECK(cb_expr_compile(dev, CB_F32, CB_KERNEL_APPLY, progs, counts, 1, &f));

    float *x = malloc(N * sizeof(float)), *y = malloc(N * sizeof(float));
    for (int i = 0; i < N; i++) x[i] = (float)i * 1e-4f - 5.0f;
    uint64_t dx = 0, dy = 0, dsum = 0;
    CHECK(cb_alloc(dev, N * sizeof(float), 0, &dx));
    CHECK(cb_alloc(dev, N * sizeof(float), 1, &dy));
    CHECK(cb_alloc(dev, 64, 1, &dsum));
    CHECK(cb_h2d(dev, dx, x, N * sizeof(float)));
    CHECK(cb_apply(dev, f, dx, dy, N));
    CHECK(cb_d2h(dev, y, dy, N * sizeof(float)));
    double worst = 0.0;
    for (int i = 0; i < N; i++) {
        const float t = x[i] * 2.0f + 1.0f; /* two roundings, like the CPU device */
        const double d = fabs((double)y[i] - (double)sinf(t));
        if (d > worst) worst = d;
    }
    printf("max |gpu - libm| = %.3e over %d elements\n", worst, N);
    if (worst > 5e-7) return 1;

    /* binary add, bit exact; then the deterministic sum */
    CHECK(cb_binary(dev, CB_F32, CB_BIN_ADD, dx, dy, dy, N));
    float s = 0.0f;
    CHECK(cb_sum_host(dev, CB_F32, dy, N, &s));
    float *z = malloc(N * sizeof(float));
    CHECK(cb_d2h(dev, z, dy, N * sizeof(float)));
    double truth = 0.0;
    for (int i = 0; i < N; i++) {
        if (z[i] != x[i] + y[i]) {
            fprintf(stderr, "binary add differs at %d\n", i);
            return 1;
        }
        truth += (double)z[i];
    }
    printf("sum = %.6f (fp64 reference %.6f)\n", s, truth);
    if (fabs((double)s - truth) > 1e-6 * 3.0e5) return 1;

    uint64_t launches = 0;
    CHECK(cb_launch_count(dev, &launches));
    printf("kernels launched: %llu\n", (unsigned long long)launches);
    CHECK(cb_free(dev, dx));
    CHECK(cb_free(dev, dy));
    CHECK(cb_free(dev, dsum));
    CHECK(cb_device_destroy(dev));
    free(x);
    free(y);
    free(z);
    puts("c_abi_smoke ok");
    return 0;
}
