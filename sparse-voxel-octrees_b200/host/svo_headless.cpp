// svo_headless -- the SDL-free replacement for the reference's `-viewer` mode
// (reference src/Main.cpp:332-376): loads an .oct, renders an orbit of frames on
// the GPU with the reference's strip/tile/beam semantics and writes PPM images.
//
//   svo_headless <in.oct> [--size WxH] [--strips N] [--frames K] [--radius R] [--pitch P]
//                [--yaw0 Y] [--yaw-step S] [--validation] [--preview] [--events script]
//                [--out prefix] [--png prefix] [--raw file|-] [--check] [--gpus N]
//
// --gpus N (N > 1) renders on devices 0..N-1 of this node from this one process (svo_multi_*: the node array is
// replicated, tile columns are dealt to the devices, every device ships its own pixels to the host frame; the
// reference's strip threads and frame barrier, Main.cpp:351-367, :217-219); the orbit is rendered as ONE pipelined
// sequence (four frames in flight) and every finished frame is written by a callback.
// --check decodes the file and walks its whole node array on the host (svo_words_validate: every pointer in bounds, no
// cycles, no branch deeper than the kernels' stack), prints the counts and exits; no GPU is needed.
// --out writes prefix_<k>.ppm, --png prefix_<k>.png (8-bit RGB, host/png_write.hpp) for every frame.
//
// --raw streams every frame as packed RGB24 (row-major, no header) to a file or to stdout ("-"), the
// input format of `ffmpeg -f rawvideo -pix_fmt rgb24 -s WxH -i -`: the streamed stand-in for the
// reference's interactive SDL window; status text goes to stderr then.
// --events replays a script of window events through the reference viewer's camera control
// (svo_viewer_feed: Events.cpp's mouse state + renderLoop's event handling, Main.cpp:229-252): one frame at
// the start and one after every event the reference would redraw for, in preview resolution (stride 3) while
// a drag is going on -- the frames the reference's window shows for the same input. Script lines:
//   motion <dx> <dy> | down left|right|<n> | up left|right|<n> | key esc|<code> | keyup esc|<code> | # comment
//   svo_headless -builder [--resolution r --mode m] <in.ply | in.voxel> <out.oct>
//
// The second form is the reference's `-builder` mode with its own argument layout (reference
// src/Main.cpp:291-330): a PLY mesh is voxelised and the tree built on the GPU (PlyLoader + VoxelData +
// VoxelOctree, :320-325; the mode flag is accepted and ignored: nothing goes through the disk), a raw
// .voxel volume takes the on-disk path's second half (VoxelData(path) + VoxelOctree, :318-319); then save.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "VoxelOctree.hpp"
#include "png_write.hpp"

static void writePpm(const std::string &path, const uint32_t *rgba, int w, int h) {
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) { fprintf(stderr, "cannot write %s\n", path.c_str()); return; }
    fprintf(fp, "P6\n%d %d\n255\n", w, h);
    std::vector<unsigned char> row(size_t(w)*3);
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            uint32_t p = rgba[size_t(y)*w + x];
            row[3*x] = p & 255; row[3*x + 1] = (p >> 8) & 255; row[3*x + 2] = (p >> 16) & 255;
        }
        fwrite(row.data(), 1, row.size(), fp);
    }
    fclose(fp);
}

struct FrameSink {      // what happens to a finished host frame: PPM / PNG files, the raw RGB24 stream
    std::string out, png;
    FILE *rawFile = 0;
    std::vector<unsigned char> rgb;
    int w = 0, h = 0;
    void write(int k, const uint32_t *rgba) {
        if (!out.empty()) writePpm(out + "_" + std::to_string(k) + ".ppm", rgba, w, h);
        if (!png.empty() && !svo_png::writeRgb(png + "_" + std::to_string(k) + ".png", rgba, w, h))
            throw std::runtime_error("cannot write " + png + "_" + std::to_string(k) + ".png");
        if (rawFile) {
            for (size_t p = 0; p < size_t(w)*h; ++p) {
                rgb[3*p] = rgba[p] & 255; rgb[3*p + 1] = (rgba[p] >> 8) & 255; rgb[3*p + 2] = (rgba[p] >> 16) & 255;
            }
            if (fwrite(rgb.data(), 1, rgb.size(), rawFile) != rgb.size()) throw std::runtime_error("short write on the raw stream");
        }
    }
};

static void sinkCallback(void *user, int frame, const uint32_t *rgba) { static_cast<FrameSink *>(user)->write(frame, rgba); }

static std::vector<svo_viewer_event> readEventScript(const std::string &path) {
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) throw std::runtime_error("cannot read " + path);
    std::vector<svo_viewer_event> out;
    char line[256];
    int lineNo = 0;
    auto code = [](const char *w) { return !strcmp(w, "left") ? int(SVO_BUTTON_LEFT) : !strcmp(w, "right") ? int(SVO_BUTTON_RIGHT) :
                                           !strcmp(w, "esc") ? int(SVO_KEY_ESCAPE) : atoi(w); };
    while (fgets(line, sizeof line, fp)) {
        ++lineNo;
        char kind[32] = "", arg[32] = "";
        int dx = 0, dy = 0;
        svo_viewer_event e = {0, 0, 0, 0};
        if (sscanf(line, " %31s", kind) != 1 || kind[0] == '#') continue;
        if (!strcmp(kind, "motion") && sscanf(line, " %*s %d %d", &dx, &dy) == 2) { e.type = SVO_EVENT_MOUSE_MOTION; e.dx = dx; e.dy = dy; }
        else if (!strcmp(kind, "down") && sscanf(line, " %*s %31s", arg) == 1) { e.type = SVO_EVENT_BUTTON_DOWN; e.code = code(arg); }
        else if (!strcmp(kind, "up") && sscanf(line, " %*s %31s", arg) == 1) { e.type = SVO_EVENT_BUTTON_UP; e.code = code(arg); }
        else if (!strcmp(kind, "key") && sscanf(line, " %*s %31s", arg) == 1) { e.type = SVO_EVENT_KEY_DOWN; e.code = code(arg); }
        else if (!strcmp(kind, "keyup") && sscanf(line, " %*s %31s", arg) == 1) { e.type = SVO_EVENT_KEY_UP; e.code = code(arg); }
        else { fclose(fp); throw std::runtime_error(path + ":" + std::to_string(lineNo) + ": cannot parse event: " + line); }
        out.push_back(e);
    }
    fclose(fp);
    return out;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <in.oct> [--size WxH] [--strips N] [--frames K] [--radius R] [--pitch P] [--yaw0 Y] "
                        "[--yaw-step S] [--validation] [--preview] [--events script] [--out prefix] [--png prefix] [--raw file|-] [--check] [--gpus N]\n"
                        "       %s -builder [--resolution r --mode m] <in.ply | in.voxel> <out.oct>\n", argv[0], argv[0]);
        return 2;
    }
    if (std::string(argv[1]) == "-builder") {
        // the reference's argument forms (Main.cpp:291-300): -builder --resolution r --mode m in out | -builder in out
        int resolution = 256;
        const char *input = 0, *output = 0;
        if (argc == 8) { resolution = atoi(argv[3]); input = argv[6]; output = argv[7]; }
        else if (argc == 4) { input = argv[2]; output = argv[3]; }
        else { fprintf(stderr, "usage: %s -builder [--resolution r --mode m] <in.ply | in.voxel> <out.oct>\n", argv[0]); return 2; }
        try {
            auto t0 = std::chrono::steady_clock::now();
            const std::string in = input;
            const bool isPly = in.size() >= 4 && in.compare(in.size() - 4, 4, ".ply") == 0;
            VoxelOctree *tree = isPly ? VoxelOctree::fromPly(input, resolution) : VoxelOctree::fromVoxelFile(input);
            svo_build_stats st;
            svo_build_last_stats(&st);
            tree->save(output);
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("built %s: %llu voxels -> %llu words (depth %u, %llu far blocks), device %.3f ms; "
                   "Octree initialization took %.3f s\n", output, (unsigned long long)st.voxels,
                   (unsigned long long)tree->wordCount(), tree->depth(), (unsigned long long)st.far_blocks,
                   st.gather_ms + st.sort_ms + st.levels_ms + st.emit_ms, s);
            delete tree;
        } catch (const std::exception &e) {
            fprintf(stderr, "error: %s\n", e.what());
            return 1;
        }
        return 0;
    }
    int w = 1280, h = 720, strips = 16, frames = 1, flavour = SVO_FLAVOUR_FAST; /* Main.cpp:57-60 defaults */
    float radius = 1.0f, pitch = 0.0f, yaw0 = 0.0f, yawStep = 3.6f;
    int stride = 1, gpus = 1;
    std::string out, raw, events, png;
    bool check = false;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--size") sscanf(next(), "%dx%d", &w, &h);
        else if (a == "--strips") strips = atoi(next());
        else if (a == "--frames") frames = atoi(next());
        else if (a == "--radius") radius = float(atof(next()));
        else if (a == "--pitch") pitch = float(atof(next()));
        else if (a == "--yaw0") yaw0 = float(atof(next()));
        else if (a == "--yaw-step") yawStep = float(atof(next()));
        else if (a == "--validation") flavour = SVO_FLAVOUR_VALIDATION;
        else if (a == "--preview") stride = 3;   /* the reference's renderHalfSize while dragging, Main.cpp:161 */
        else if (a == "--out") out = next();
        else if (a == "--raw") raw = next();
        else if (a == "--png") png = next();
        else if (a == "--check") check = true;
        else if (a == "--gpus") gpus = atoi(next());
        else if (a == "--events") events = next();
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (check) {
        // host only: decode the file and walk the whole node array (svo_words_validate); no GPU is touched
        uint32_t *words = 0;
        uint64_t n = 0;
        float center[3];
        svo_words_report rep;
        if (svo_oct_read(argv[1], &words, &n, center) != SVO_OK || svo_words_validate(words, n, &rep) != SVO_OK) {
            fprintf(stderr, "error: %s: %s\n", argv[1], svo_last_error());
            svo_free(words);
            return 1;
        }
        printf("%s: ok -- %llu words = %llu descriptors + %llu far words + %llu leaf words%s, depth %u, centre (%g, %g, %g)\n", argv[1],
               (unsigned long long)n, (unsigned long long)rep.descriptors, (unsigned long long)rep.far_words,
               (unsigned long long)rep.leaves, rep.descriptors + rep.far_words + rep.leaves == n ? "" : " (+ unreachable words)",
               rep.depth, center[0], center[1], center[2]);
        svo_free(words);
        return 0;
    }
    try {
        std::vector<svo_viewer_event> script;
        if (!events.empty()) script = readEventScript(events);     // before anything touches the GPU: a bad script fails fast
        if (gpus > 1) {
            // several GPUs, one process: the whole orbit (or every frame of the replayed session) through svo_multi_*
            FILE *info = raw == "-" ? stderr : stdout;
            FrameSink sink;
            sink.out = out; sink.png = png; sink.w = w; sink.h = h;
            sink.rawFile = raw.empty() ? 0 : raw == "-" ? stdout : fopen(raw.c_str(), "wb");
            if (!raw.empty() && !sink.rawFile) throw std::runtime_error("cannot write " + raw);
            sink.rgb.resize(raw.empty() ? 0 : size_t(w)*h*3);
            std::vector<int> devices;
            for (int d = 0; d < gpus; ++d) devices.push_back(d);
            svo_multi *multi = 0;
            if (svo_multi_load_oct(argv[1], devices.data(), gpus, &multi) != SVO_OK) throw std::runtime_error(svo_last_error());
            svo_tree_info ti;
            svo_tree_get_info(svo_multi_tree(multi, 0), &ti);
            fprintf(info, "loaded %s on %d GPUs: %llu words, depth %u\n", argv[1], gpus, (unsigned long long)ti.n_words, ti.depth);
            std::vector<svo_camera> path;
            std::vector<int> strideOf;
            if (events.empty()) {
                for (int k = 0; k < frames; ++k) { svo_camera c; svo_orbit_camera(pitch, yaw0 + yawStep*k, radius, &c); path.push_back(c); strideOf.push_back(stride); }
            } else {
                svo_viewer_state viewer;
                svo_viewer_init(&viewer);
                path.push_back(viewer.camera); strideOf.push_back(1);
                for (size_t e = 0; e < script.size();) {
                    int action = SVO_VIEWER_WAIT;
                    while (action == SVO_VIEWER_WAIT && e < script.size()) action = svo_viewer_feed(&viewer, &script[e++]);
                    if (action != SVO_VIEWER_FRAME) break;
                    path.push_back(viewer.camera); strideOf.push_back(viewer.preview ? 3 : 1);
                }
            }
            uint32_t *ring[4] = {0, 0, 0, 0};
            for (int i = 0; i < 4; ++i)
                if (svo_host_alloc(size_t(w)*h*4, reinterpret_cast<void **>(&ring[i])) != SVO_OK) throw std::runtime_error(svo_last_error());
            double totalMs = 0.0;
            unsigned long long rays = 0;
            // runs of frames with the same pixel stride go out as one pipelined sequence each
            for (size_t first = 0; first < path.size();) {
                size_t last = first;
                while (last < path.size() && strideOf[last] == strideOf[first]) ++last;
                svo_frame_desc desc = {w, h, strips, flavour, 0, 1, strideOf[first], 0};
                svo_sequence_stats st;
                struct Shifted { FrameSink *sink; int base; } shifted = {&sink, int(first)};
                auto cb = [](void *user, int frame, const uint32_t *rgba) { Shifted *s = static_cast<Shifted *>(user); sinkCallback(s->sink, s->base + frame, rgba); };
                if (svo_multi_render_sequence(multi, path.data() + first, int(last - first), &desc, SVO_OUTPUT_HOST, ring, 4, cb, &shifted, &st) != SVO_OK)
                    throw std::runtime_error(svo_last_error());
                totalMs += st.wall_ms;
                rays += st.coarse_rays + st.fine_rays;
                first = last;
            }
            fprintf(info, "%zu frame(s) %dx%d, %d strips, %d GPUs: %.3f ms/frame end to end (host buffers, pipelined), %.1f Mrays/s\n",
                    path.size(), w, h, strips, gpus, totalMs/double(path.size()), double(rays)/(totalMs*1e3));
            if (sink.rawFile && sink.rawFile != stdout) fclose(sink.rawFile);
            for (int i = 0; i < 4; ++i) svo_host_free(ring[i]);
            svo_multi_destroy(multi);
            return 0;
        }
        VoxelOctree tree(argv[1]);
        FILE *info = raw == "-" ? stderr : stdout;
        FILE *rawFile = raw.empty() ? 0 : raw == "-" ? stdout : fopen(raw.c_str(), "wb");
        if (!raw.empty() && !rawFile) throw std::runtime_error("cannot write " + raw);
        std::vector<unsigned char> rgb(raw.empty() ? 0 : size_t(w)*h*3);
        fprintf(info, "loaded %s: %llu words, depth %u\n", argv[1], (unsigned long long)tree.wordCount(), tree.depth());
        uint32_t *rgba = 0;
        if (svo_host_alloc(size_t(w)*h*4, reinterpret_cast<void **>(&rgba)) != SVO_OK) throw std::runtime_error(svo_last_error());
        double totalMs = 0.0;
        unsigned long long rays = 0;
        svo_viewer_state viewer;
        svo_viewer_init(&viewer);
        size_t nextEvent = 0;
        if (!events.empty()) frames = int(script.size()) + 1;      // at most one frame per event, plus the first
        for (int k = 0; k < frames; ++k) {
            svo_camera cam;
            if (events.empty()) {
                svo_orbit_camera(pitch, yaw0 + yawStep*k, radius, &cam);
            } else {
                if (k > 0) {                                       // wait for an event that redraws (Main.cpp:229)
                    int action = SVO_VIEWER_WAIT;
                    while (action == SVO_VIEWER_WAIT && nextEvent < script.size()) action = svo_viewer_feed(&viewer, &script[nextEvent++]);
                    if (action != SVO_VIEWER_FRAME) { frames = k; break; }   // Escape, or the script ran out
                }
                cam = viewer.camera;
                stride = viewer.preview ? 3 : 1;
            }
            auto t0 = std::chrono::steady_clock::now();
            svo_frame_stats st = tree.renderFrame(cam, w, h, strips, rgba, flavour, stride);
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (k > 0 || frames == 1) { totalMs += ms; rays += st.coarse_rays + st.fine_rays; }
            if (!out.empty()) writePpm(out + "_" + std::to_string(k) + ".ppm", rgba, w, h);
            if (!png.empty() && !svo_png::writeRgb(png + "_" + std::to_string(k) + ".png", rgba, w, h))
                throw std::runtime_error("cannot write " + png + "_" + std::to_string(k) + ".png");
            if (rawFile) {
                for (size_t p = 0; p < size_t(w)*h; ++p) {
                    rgb[3*p] = rgba[p] & 255; rgb[3*p + 1] = (rgba[p] >> 8) & 255; rgb[3*p + 2] = (rgba[p] >> 16) & 255;
                }
                if (fwrite(rgb.data(), 1, rgb.size(), rawFile) != rgb.size()) throw std::runtime_error("short write on the raw stream");
            }
        }
        int timed = frames > 1 ? frames - 1 : 1;
        fprintf(info, "%d frame(s) %dx%d, %d strips: %.3f ms/frame end to end (host buffer), %.1f Mrays/s\n", timed, w, h,
                strips, totalMs/timed, rays/(totalMs*1e3));
        if (rawFile && rawFile != stdout) fclose(rawFile);
        svo_host_free(rgba);
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
