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

    /* ---- the north-star stack from C: CUDA<Lazy<Graph<Autograd<Base>>>>, three unary_ew ops (x * 2, sin, + 1),
     * unary_fusing, run() and backward(): ONE forward kernel and ONE recomputing backward kernel ---------------- */
    {
        cbm_device *md = NULL;
        CHECK(cbm_device_create(0, CBM_LAZY | CBM_GRAPH | CBM_AUTOGRAD, CB_F32, &md));
        cbm_buf bx = 0, b1 = 0, b2 = 0, b3 = 0, bg = 0;
        CHECK(cbm_buffer_from_host(md, CB_F32, x, N, &bx));
        CHECK(cbm_buffer_require_grad(md, bx));
        const cb_node f_mul2[3] = {{CB_OP_X, -1, -1, 0, 0.0, 0}, {CB_OP_CONST, -1, -1, 0, 2.0, 0}, {CB_OP_MUL, 0, 1, 0, 0.0, 0}};
        const cb_node g_two[1] = {{CB_OP_CONST, -1, -1, 0, 2.0, 0}};
        const cb_node f_sin[2] = {{CB_OP_X, -1, -1, 0, 0.0, 0}, {CB_OP_SIN, 0, -1, 0, 0.0, 0}};
        const cb_node g_cos[2] = {{CB_OP_X, -1, -1, 0, 0.0, 0}, {CB_OP_COS, 0, -1, 0, 0.0, 0}};
        const cb_node f_add1[3] = {{CB_OP_X, -1, -1, 0, 0.0, 0}, {CB_OP_CONST, -1, -1, 0, 1.0, 0}, {CB_OP_ADD, 0, 1, 0, 0.0, 0}};
        const cb_node g_one[1] = {{CB_OP_CONST, -1, -1, 0, 1.0, 0}};
        CHECK(cbm_unary_ew(md, bx, f_mul2, 3, g_two, 1, &b1));
        CHECK(cbm_unary_ew(md, b1, f_sin, 2, g_cos, 2, &b2));
        CHECK(cbm_unary_ew(md, b2, f_add1, 3, g_one, 1, &b3));
        CHECK(cbm_optimize_mem_graph(md));
        CHECK(cbm_unary_fusing(md));
        CHECK(cbm_run(md));
        CHECK(cbm_backward(md, b3));
        cb_device *raw = NULL;
        CHECK(cbm_device_raw(md, &raw));
        uint64_t before = 0, after = 0;
        CHECK(cb_launch_count(raw, &before));
        CHECK(cbm_run(md));
        CHECK(cbm_backward(md, b3));
        CHECK(cb_launch_count(raw, &after));
        printf("fused forward + backward: %llu kernels\n", (unsigned long long)(after - before));
        if (after - before != 2) return 1;
        CHECK(cbm_buffer_read(md, b3, y, N));
        CHECK(cbm_grad(md, bx, &bg));
        CHECK(cbm_buffer_read(md, bg, z, N));
        double worst_y = 0.0, worst_g = 0.0;
        for (int i = 0; i < N; i++) {
            const float t = x[i] * 2.0f;
            const double dy_ = fabs((double)y[i] - (double)(sinf(t) + 1.0f));
            const double dg = fabs((double)z[i] / 2.0 - (double)(cosf(t) * 2.0f)); /* two backward passes accumulated */
            if (dy_ > worst_y) worst_y = dy_;
            if (dg > worst_g) worst_g = dg;
        }
        printf("module stack: max |y - libm| = %.3e, max |grad - libm| = %.3e\n", worst_y, worst_g);
        if (worst_y > 5e-7 || worst_g > 1e-6) return 1;
        CHECK(cbm_device_destroy(md));
    }

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
