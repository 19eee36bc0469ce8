// Exercises the drop-in C++ facade (sparse-voxel-octrees_b200/host/VoxelOctree.hpp) the way the reference's
// own code uses its VoxelOctree (src/Main.cpp:118,181,333; src/VoxelOctree.hpp:48-57): construct from a
// path, center(), raymarch() with untouched-output semantics, save(), adopt(). Prints one line per check;
// the Python test compares the numbers with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "VoxelOctree.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    try {
        VoxelOctree tree(argv[1]);
        Vec3 c = tree.center();
        printf("center %.9g %.9g %.9g words %llu depth %u\n", c.x, c.y, c.z, (unsigned long long)tree.wordCount(), tree.depth());

        // the reference's eye position for the default camera: tform*0 + center + 1 with VIEW = translate(0,0,-1)
        Vec3 o(c.x + 1.0f, c.y + 1.0f, c.z + 1.0f - 1.0f);
        const float dirs[4][3] = {{0.0f, 0.0f, 1.0f}, {0.0f, 1.0f, 0.0f}, {0.05f, -0.02f, 1.0f}, {-0.3f, 0.1f, 1.0f}};
        const float scales[4] = {0.0f, 0.0f, 0.05f, 0.0f};
        for (int i = 0; i < 4; ++i) {
            uint32 normal = 0xABCD1234u;
            float t = -7.0f;
            bool hit = tree.raymarch(o, Vec3(dirs[i][0], dirs[i][1], dirs[i][2]), scales[i], normal, t);
            unsigned tbits;
            memcpy(&tbits, &t, 4);
            printf("ray %d hit %d normal %08x tbits %08x\n", i, hit ? 1 : 0, normal, tbits);
        }

        tree.save(argv[2]);
        VoxelOctree again(argv[2]);
        printf("reload words %llu\n", (unsigned long long)again.wordCount());

        // adopt(): a node array handed over by a builder
        uint32_t *words = 0;
        uint64_t n = 0;
        float center[3];
        if (svo_oct_read(argv[1], &words, &n, center) != SVO_OK) throw std::runtime_error(svo_last_error());
        VoxelOctree *adopted = VoxelOctree::adopt(words, n, Vec3(center[0], center[1], center[2]));
        svo_free(words);
        svo_camera cam;
        svo_orbit_camera(20.0f, 135.0f, 0.7f, &cam);
        std::vector<uint32_t> rgba(size_t(160)*90);
        svo_frame_stats st = adopted->renderFrame(cam, 160, 90, 4, rgba.data(), SVO_FLAVOUR_VALIDATION);
        unsigned long long sum = 0;
        for (size_t i = 0; i < rgba.size(); ++i) sum = sum*1099511628211ull + rgba[i];
        printf("frame rays %llu fnv %016llx\n", (unsigned long long)(st.coarse_rays + st.fine_rays), sum);
        delete adopted;

        // the same four rays through the queue: one batch, the reference's untouched-output semantics
        VoxelOctree::RayQueue queue;
        for (int i = 0; i < 4; ++i) queue.push(o, Vec3(dirs[i][0], dirs[i][1], dirs[i][2]));
        queue.flush(tree, 0.0f);
        for (int i = 0; i < 4; ++i) {
            uint32 normal = 0xABCD1234u;
            float t = -7.0f;
            bool hit = queue.result(size_t(i), normal, t);
            unsigned tbits;
            memcpy(&tbits, &t, 4);
            printf("queued %d hit %d normal %08x tbits %08x\n", i, hit ? 1 : 0, normal, tbits);
        }

        // VoxelOctree(VoxelData*) with the reference's constructor forms (Main.cpp:318-319): a raw .voxel file
        if (argc > 3) {
            VoxelData data(argv[3], size_t(1) << 30);
            VoxelOctree built(&data);
            printf("built words %llu depth %u\n", (unsigned long long)built.wordCount(), built.depth());
        }

        try {
            VoxelOctree missing("/nonexistent/file.oct");
            printf("missing: no error\n");
        } catch (const std::exception &e) {
            printf("missing: threw\n");
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
