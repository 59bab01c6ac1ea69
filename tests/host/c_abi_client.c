/* A plain-C client of the boundary (include/skyb200.h + libskyb200.so): no C++, no Python, no torch types.
 * Host-only entry points and argument validation, so it runs without a GPU (tests/test_abi_cpu.py). */
#include <stdio.h>

#include "skyb200.h"

int main(void) {
    int64_t plan[3];
    if (skyb200_abi_version() != 2) return 1;
    if (skyb200_policy_step_plan(256, 4096, 0, 3584, 152064, SKYB200_LOSS_REGULAR, 1, plan) != 0) return 2;
    size_t ws = skyb200_policy_step_workspace(256, 4096, 0, 3584, 152064, SKYB200_LOSS_REGULAR, 1);
    /* argument validation happens before any CUDA call: a NULL pointer is an error code plus a message */
    int rc = skyb200_gae_advantage(0, 0, 0, SKYB200_MASK_F32, 0, 0, 0, 0, 0, 4, 4, 1.0f, 0.95f, 0);
    printf("%lld %lld %lld %zu %d %s\n", (long long)plan[0], (long long)plan[1], (long long)plan[2], ws, rc,
           skyb200_last_error());
    return 0;
}
