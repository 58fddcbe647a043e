/* Minimal C client of libgnnb.so: proves that include/gnnb.h is usable from plain C and that the library refuses to run
 * without a B200 (no CPU fallback).  Exit code 0 = behaved as expected; prints what it saw.
 *   gcc -std=c99 -I include tests/c/abi_client.c -L gnn_branching_b200 -lgnnb -o abi_client */
#include <stdio.h>
#include "gnnb.h"

int main(void) {
    gnnb_ctx* ctx = NULL;
    int v = gnnb_abi_version();
    int st = gnnb_create(&ctx, 0);
    printf("abi %d create %d ctx %s\n", v, st, ctx ? "set" : "null");
    if (v != 1) return 2;
    if (st == GNNB_OK) {                       /* a B200 is present: the context must be usable and destroyable */
        if (!ctx) return 3;
        if (gnnb_get_option(ctx, "math") != GNNB_MATH_TC_FP16X3) return 4;
        if (gnnb_score(ctx, NULL, NULL, NULL, NULL, NULL) != GNNB_ERR_INVALID) return 5;
        gnnb_destroy(ctx);
        return 0;
    }
    /* no usable device: a status code, no context, never a silent CPU path */
    if (ctx != NULL) return 6;
    if (st != GNNB_ERR_CUDA && st != GNNB_ERR_UNSUPPORTED) return 7;
    return 0;
}
