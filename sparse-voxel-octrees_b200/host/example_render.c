/* The C ABI from plain C99: load an .oct, render one frame with the reference's strip / tile / beam semantics, cast one
 * ray with VoxelOctree::raymarch's contract, write a PPM.   cc -std=c99 -I../../include example_render.c -L.. -lsvo_b200
 *
 *   example_render <in.oct> <out.ppm> [width height]
 */
#include <stdio.h>
#include <stdlib.h>

#include "svo_b200.h"

static int fail(const char *what) {
    fprintf(stderr, "%s: %s\n", what, svo_last_error());
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <in.oct> <out.ppm> [width height]\n", argv[0]);
        return 2;
    }
    const int width = argc > 4 ? atoi(argv[3]) : 1280, height = argc > 4 ? atoi(argv[4]) : 720;   /* Main.cpp:59-60 */
    svo_tree *tree = NULL;
    if (svo_tree_load_oct(argv[1], 0, &tree) != SVO_OK) return fail("svo_tree_load_oct");        /* VoxelOctree(path) */

    svo_tree_info info;
    if (svo_tree_get_info(tree, &info) != SVO_OK) return fail("svo_tree_get_info");
    printf("%llu words, depth %u, centre (%g, %g, %g)\n", (unsigned long long)info.n_words, info.depth,
           info.center[0], info.center[1], info.center[2]);

    /* one ray from the default eye position along +z: bool raymarch(o, d, rayScale, normal&, t&), VoxelOctree.cpp:207 */
    float o[3] = {info.center[0] + 1.0f, info.center[1] + 1.0f, info.center[2]}, d[3] = {0.0f, 0.0f, 1.0f}, t = 0.0f;
    uint32_t normal = 0;
    int hit = 0;
    if (svo_raymarch(tree, o, d, 0.0f, &normal, &t, &hit) != SVO_OK) return fail("svo_raymarch");
    printf("centre ray: %s", hit ? "hit" : "miss");
    if (hit) printf(" at t = %g, material word %08x", t, normal);
    printf("\n");

    /* one frame: the viewer's first camera (Main.cpp:207-213), 16 strips like the reference's 16 threads */
    svo_viewer_state viewer;
    svo_viewer_init(&viewer);
    svo_frame_desc desc = {0};
    desc.width = width; desc.height = height; desc.strips = 16;
    desc.flavour = SVO_FLAVOUR_FAST; desc.tile_rank = 0; desc.tile_world = 1; desc.pixel_stride = 1;
    uint32_t *pixels = NULL;
    if (svo_host_alloc((size_t)width*(size_t)height*4, (void **)&pixels) != SVO_OK) return fail("svo_host_alloc");
    svo_frame_stats stats;
    if (svo_render_frame(tree, &viewer.camera, &desc, pixels, NULL, &stats) != SVO_OK) return fail("svo_render_frame");
    printf("%llu beam rays + %llu pixel rays, %llu of %llu tiles rendered\n", (unsigned long long)stats.coarse_rays,
           (unsigned long long)stats.fine_rays, (unsigned long long)stats.tiles_rendered, (unsigned long long)stats.tiles_total);

    FILE *fp = fopen(argv[2], "wb");
    if (!fp) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    fprintf(fp, "P6\n%d %d\n255\n", width, height);
    for (size_t p = 0; p < (size_t)width*(size_t)height; ++p) {       /* backBuffer words: 0xFF000000 | b << 16 | g << 8 | r */
        unsigned char rgb[3] = {(unsigned char)pixels[p], (unsigned char)(pixels[p] >> 8), (unsigned char)(pixels[p] >> 16)};
        fwrite(rgb, 1, 3, fp);
    }
    fclose(fp);
    svo_host_free(pixels);
    svo_tree_destroy(tree);
    return 0;
}
