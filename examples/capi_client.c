/* Minimal C client of the drop-in boundary (include/lidarnerf_b200.h -> liblnb200.so): no Python, no torch, plain
 * pointers and sizes.  Without a GPU it only exercises the argument checks (every entry point validates before it
 * launches); on a GPU box pass device pointers obtained from cudaMalloc exactly as the reference's bindings pass
 * tensor.data_ptr().
 *
 *   gcc -std=c99 -Iinclude examples/capi_client.c -Llidar-nerf_b200/lib -llnb200 -Wl,-rpath,$PWD/lidar-nerf_b200/lib -o /tmp/capi_client
 */
#include <stdio.h>
#include <string.h>

#include "lidarnerf_b200.h"

int main(void) {
    printf("liblnb200 built for %s\n", lnb_arch());
    if (strcmp(lnb_arch(), "sm_100a") != 0) return 1;
    /* NULL buffers are rejected with a negative LNB_ERR_* status before any launch */
    int rc = lnb_march_rays_train(NULL, NULL, NULL, 1.0f, 0.0f, 1024, 16, 1, 128, 1024, NULL, NULL, NULL, NULL, NULL, NULL,
                                  NULL, NULL, NULL);
    printf("march_rays_train(NULL...) -> %d (%s)\n", rc, lnb_strerror(rc));
    if (rc >= 0) return 2;
    rc = lnb_chamfer_forward(NULL, NULL, 1, 8, 8, NULL, NULL, NULL, NULL, NULL);
    printf("chamfer_forward(NULL...)  -> %d (%s)\n", rc, lnb_strerror(rc));
    if (rc >= 0) return 3;
    printf("workspace for a 64x1024 range image: %zu bytes\n", lnb_lidar_to_pano_workspace_bytes(64, 1024));
    if (lnb_lidar_to_pano_workspace_bytes(64, 1024) != (size_t)64 * 1024 * 8) return 4;
    printf("kernel launches so far: %llu\n", (unsigned long long)lnb_launch_count());
    return 0;
}
