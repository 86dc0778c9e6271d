// stereo_vision_parallel -- the sequence driver, re-implemented without OpenCV / popt / GLUT.
//
// Replaces main(), imageLoop() and runProfiling() of src/parallel_includes/main/stereo_vision.cu:645-858:
//   stereo_vision_parallel -k <kitti sequence> [-v N] [-p 0] [-s 0] [-f 1.0] [-w 1242] [-h 375] [-e 1] [-t 0] [-d 0]
//       reads <k>/image_02/data/%010u.png and <k>/image_03/data/%010u.png (:658-659), calibration from the cwd-relative
//       data/calibration/kitti_2011_09_26.yml (:66), prints one "(FPS=...) (rows, cols) (t_t=..., dmap_t=..., pc_t=...)"
//       line per frame (:691) and "AVG_FPS=" at the end (:695) -- the lines test.py scrapes.
//   stereo_vision_parallel -P 1
//       runProfiling(): the seven PGM pairs under datasets/profile, ROBOTICS parameters with both maps post-processed,
//       writes <name>_disp.pgm scaled by the per-pair maximum (:699-764).
// Option letters, long names and the "-x=val" / "-x val" / "--name=val" spellings follow the popt table of :768-784.
// Extension (not in the reference): -B N processes the sequence in batches of N frames through the frame-batch
// pipeline and reports the aggregate rate.
// Out of scope here and ignored / rejected with a message: -p 1 (OpenGL viewer), -t 1 (YOLO object tracking), -e != 1.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <chrono>
#include <future>
#include <string>
#include <vector>

#include "../../include/elas.h"
#include "../../include/elas_b200.h"
#include "../../include/stereo_vision_c.h"
#include "image_io.h"

namespace {

struct Options {
    std::string kitti_path;
    int subsample = 0, video_mode = 0, draw_points = 1, debug = 0, object_tracking = 0;
    int width = 1242, height = 375, extrapolate = 1, profile = 0, batch = 0;
    std::string dump_dir;
    float scale_factor = 1.f;
    std::string calib = "data/calibration/kitti_2011_09_26.yml";
};

struct OptDef {
    const char *long_name;
    char short_name;
    char kind;  // 's' string, 'i' int, 'f' float
    void *target;
    const char *help;
};

void usage(const std::vector<OptDef> &defs) {
    fprintf(stderr, "Usage: stereo_vision_parallel");
    for (const auto &d : defs) fprintf(stderr, " [-%c|--%s=%s]", d.short_name, d.long_name, d.kind == 's' ? "STR" : "NUM");
    fprintf(stderr, "\n");
    for (const auto &d : defs) fprintf(stderr, "  -%c, --%-26s %s\n", d.short_name, d.long_name, d.help);
}

bool assign(const OptDef &d, const char *val) {
    if (!val) return false;
    if (d.kind == 's') *(std::string *)d.target = val;
    else if (d.kind == 'i') *(int *)d.target = atoi(val);
    else *(float *)d.target = (float)atof(val);
    return true;
}

bool parse(int argc, const char **argv, const std::vector<OptDef> &defs) {
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        const OptDef *def = nullptr;
        const char *val = nullptr;
        if (a[0] == '-' && a[1] == '-') {
            const char *eq = strchr(a, '=');
            const std::string name = eq ? std::string(a + 2, eq) : std::string(a + 2);
            if (name == "help") return false;
            for (const auto &d : defs)
                if (name == d.long_name) def = &d;
            if (eq) val = eq + 1;
        } else if (a[0] == '-' && a[1]) {
            if (a[1] == '?') return false;
            for (const auto &d : defs)
                if (a[1] == d.short_name) def = &d;
            if (a[2] == '=') val = a + 3;
            else if (a[2]) val = a + 2;
        }
        if (!def) {
            fprintf(stderr, "stereo_vision: unknown option -- '%s'\n", a);
            return false;
        }
        if (!val && i + 1 < argc) val = argv[++i];
        if (!assign(*def, val)) {
            fprintf(stderr, "stereo_vision: missing argument -- '%s'\n", a);
            return false;
        }
    }
    return true;
}

bool file_exists(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

std::string frame_path(const std::string &root, const char *cam, unsigned i) {
    char buf[64];
    snprintf(buf, sizeof(buf), "/%s/data/%010u.png", cam, i);
    return root + buf;
}

bool load_bgra(const std::string &path, int W, int H, std::vector<uint8_t> *bgra) {
    svb::ImageU8 im;
    std::string err;
    if (!svb::read_png(path, &im, &err)) {
        fprintf(stderr, "%s\n", err.c_str());
        return false;
    }
    if (im.width != W || im.height != H) {
        fprintf(stderr, "%s is %dx%d, expected %dx%d (resizing is not built: pass -w/-h)\n", path.c_str(), im.width, im.height, W, H);
        return false;
    }
    bgra->resize((size_t)W * H * 4);
    if (im.channels == 4) {
        memcpy(bgra->data(), im.data.data(), bgra->size());
    } else {
        for (size_t i = 0; i < (size_t)W * H; i++) {
            const uint8_t g = im.data[i];
            (*bgra)[4 * i] = (*bgra)[4 * i + 1] = (*bgra)[4 * i + 2] = g;
            (*bgra)[4 * i + 3] = 255;
        }
    }
    return true;
}

// runProfiling (stereo_vision.cu:699-764)
void run_profiling(const std::string &file_1, const std::string &file_2) {
    printf("Processing: %s, %s\n", file_1.c_str(), file_2.c_str());
    svb::ImageU8 I1, I2;
    std::string err;
    if (!svb::read_pgm(file_1, &I1, &err) || !svb::read_pgm(file_2, &I2, &err)) {
        printf("ERROR: %s\n", err.c_str());
        return;
    }
    if (I1.width != I2.width || I1.height != I2.height) {
        printf("ERROR: Images must be of same size, but\n       I1: %d x %d, I2: %d x %d\n", I1.width, I1.height, I2.width, I2.height);
        return;
    }
    const int32_t width = I1.width, height = I1.height;
    const int32_t dims[3] = {width, height, width};
    std::vector<float> D1((size_t)width * height, 0.f), D2((size_t)width * height, 0.f);
    ElasGPU::parameters param;  // ROBOTICS
    param.postprocess_only_left = false;
    ElasGPU elas(param);
    elas.process(I1.data.data(), I2.data.data(), D1.data(), D2.data(), dims);
    float disp_max = 0;
    for (size_t i = 0; i < D1.size(); i++) {
        if (D1[i] > disp_max) disp_max = D1[i];
        if (D2[i] > disp_max) disp_max = D2[i];
    }
    std::vector<uint8_t> o1(D1.size()), o2(D2.size());
    for (size_t i = 0; i < D1.size(); i++) {
        o1[i] = (uint8_t)fmax(255.0 * D1[i] / disp_max, 0.0);
        o2[i] = (uint8_t)fmax(255.0 * D2[i] / disp_max, 0.0);
    }
    svb::write_pgm(file_1.substr(0, file_1.size() - 4) + "_disp.pgm", o1.data(), width, height, &err);
    svb::write_pgm(file_2.substr(0, file_2.size() - 4) + "_disp.pgm", o2.data(), width, height, &err);
}

int batch_loop(const Options &o, unsigned max_files) {
    const int W = o.width, H = o.height;
    const size_t N = (size_t)W * H;
    svb_params p;
    svb_default_params(SVB_PIPELINE, &p);
    const int chunk = o.batch < 32 ? o.batch : 32;
    svb_context *ctx = svb_create(&p, W, H, chunk, -1);
    if (!ctx) {
        fprintf(stderr, "%s\n", svb_last_error());
        return 1;
    }
    svb_calibration cal;
    double Q[16];
    if (svb_calib_load_yaml(o.calib.c_str(), &cal) != SVB_OK || svb_stereo_rectify(&cal, W, H, W, H, 1.0, 0.0, 0, 0, 0, 0, Q) != SVB_OK) {
        fprintf(stderr, "%s\n", svb_last_error());
        return 1;
    }
    svb_set_calibration(ctx, Q, cal.XR, cal.XT);
    uint8_t *bl = (uint8_t *)svb_host_alloc(N * 4), *br = (uint8_t *)svb_host_alloc(N * 4);
    uint8_t *gl = (uint8_t *)svb_host_alloc(N * o.batch), *gr = (uint8_t *)svb_host_alloc(N * o.batch);
    double *pts = (double *)svb_host_alloc(N * 24 * o.batch);
    if (!bl || !br || !gl || !gr || !pts) {
        fprintf(stderr, "%s\n", svb_last_error());
        return 1;
    }
    double total_s = 0;
    unsigned done = 0;
    std::vector<uint8_t> tmp;
    for (unsigned first = 0; first < max_files; first += o.batch) {
        const int n = (int)std::min<unsigned>(o.batch, max_files - first);
        for (int i = 0; i < n; i++) {  // decode + gray conversion (input side, untimed like imread in the reference)
            if (!load_bgra(frame_path(o.kitti_path, "image_02", first + i), W, H, &tmp)) return 1;
            memcpy(bl, tmp.data(), N * 4);
            if (!load_bgra(frame_path(o.kitti_path, "image_03", first + i), W, H, &tmp)) return 1;
            memcpy(br, tmp.data(), N * 4);
            if (svb_stage_bgra_to_gray(ctx, bl, gl + (size_t)i * N) != SVB_OK || svb_stage_bgra_to_gray(ctx, br, gr + (size_t)i * N) != SVB_OK) {
                fprintf(stderr, "%s\n", svb_last_error());
                return 1;
            }
        }
        const auto t0 = std::chrono::steady_clock::now();
        if (svb_batch_run_host(ctx, gl, gr, n, SVB_OUT_POINTS, nullptr, pts) != SVB_OK) {
            fprintf(stderr, "%s\n", svb_last_error());
            return 1;
        }
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        total_s += dt;
        done += n;
        printf("(BATCH frames=%d) (%d, %d) (t=%f s, %f frames/s)\n", n, H, W, dt, n / dt);
    }
    printf("AVG_FPS=%f\n", done / total_s);
    svb_destroy(ctx);
    return 0;
}

}  // namespace

int main(int argc, const char **argv) {
    Options o;
    const std::vector<OptDef> defs = {
        {"kitti_path", 'k', 's', &o.kitti_path, "Path to KITTI Dataset"},
        {"subsampling", 's', 'i', &o.subsample, "Set s=1 for evaluating only every second pixel"},
        {"video_mode", 'v', 'i', &o.video_mode, "Set v=1 Kitti video mode"},
        {"draw_points", 'p', 'i', &o.draw_points, "Set p=1 to plot out points"},
        {"debug", 'd', 'i', &o.debug, "Set d=1 for cam to robot frame calibration"},
        {"object_tracking", 't', 'i', &o.object_tracking, "Set t=1 for enabling object tracking"},
        {"input_image_width", 'w', 'i', &o.width, "Set the input image width (default value is 1242, i.e Kitti image width)"},
        {"input_image_height", 'h', 'i', &o.height, "Set the input image height (default value is 375, i.e Kitti image height)"},
        {"scale_factor", 'f', 'f', &o.scale_factor, "All operations will be applied after shrinking the image by this factor"},
        {"extrapolate_point_cloud", 'e', 'i', &o.extrapolate, "Extrapolate the point cloud by this factor"},
        {"profile", 'P', 'i', &o.profile, "Profile"},
        {"batch", 'B', 'i', &o.batch, "(extension) process the sequence in batches of N frames"},
        {"calibration", 'c', 's', &o.calib, "(extension) calibration YAML (default data/calibration/kitti_2011_09_26.yml)"},
        {"dump_dir", 'o', 's', &o.dump_dir, "(extension) write the u8 disparity map of every frame as <dir>/%010u_disp.pgm"},
    };
    if (argc < 2 || !parse(argc, argv, defs)) {
        usage(defs);
        return 1;
    }
    if (o.profile) {
        static const char *names[7] = {"cones", "aloe", "raindeer", "urban1", "urban2", "urban3", "urban4"};
        for (const char *n : names) run_profiling(std::string("datasets/profile/") + n + "_left.pgm", std::string("datasets/profile/") + n + "_right.pgm");
        printf("... done!\n");
        return 0;
    }
    if (o.object_tracking) fprintf(stderr, "object tracking (YOLO + Bayesian tracker) is outside this program's scope; continuing without it\n");
    printf("** Object tracking disabled\n");
    printf("KITTI Path: %s \n", o.kitti_path.c_str());
    if (o.draw_points) fprintf(stderr, "the OpenGL viewer is outside this program's scope (-p 1 ignored)\n");
    if (o.extrapolate < 1 || !(o.scale_factor > 0.f)) {
        fprintf(stderr, "extrapolate_point_cloud must be >= 1 and scale_factor positive\n");
        return 1;
    }
    if (o.batch > 0 && o.extrapolate != 1) {
        fprintf(stderr, "-B (batch extension) runs with extrapolate_point_cloud 1 only\n");
        return 1;
    }
    if (o.batch > 0 && o.scale_factor != 1.f) {
        fprintf(stderr, "-B (batch extension) runs at scale_factor 1 only\n");
        return 1;
    }
    unsigned max_files = 0;
    while (file_exists(frame_path(o.kitti_path, "image_02", max_files))) max_files++;
    printf("Max files = %u\n", max_files);
    if (o.batch > 0) return batch_loop(o, max_files);

    // imageLoop() (stereo_vision.cu:645-697): one frame at a time, the body generatePointCloud() also runs
    // stereo_vision.cu:817-822: calibration size = input size, every frame is resized to out size = input / scale_factor
    const int inW = o.width, inH = o.height;
    const int W = (int)(inW / o.scale_factor), H = (int)(inH / o.scale_factor);
    const size_t N = (size_t)W * H;
    svb_params p;
    svb_default_params(SVB_PIPELINE, &p);  // generateDisparityMap()'s preset (stereo_vision.cu:315-319)
    p.subsampling = o.subsample ? 1 : 0;    // -s 1: half-resolution matching (the map fills the first quarter of the float buffer)
    printf("Post Process only left = %d, Subsampling = %d\n", p.postprocess_only_left, p.subsampling);
    svb_context *ctx = svb_create(&p, W, H, 1, -1);
    svb_calibration cal;
    double Q[16];
    if (!ctx || svb_calib_load_yaml(o.calib.c_str(), &cal) != SVB_OK ||
        svb_stereo_rectify(&cal, inW, inH, W, H, o.scale_factor, 0.0, 0, 0, 0, 0, Q) != SVB_OK) {
        fprintf(stderr, "%s\n", svb_last_error());
        return 1;
    }
    if (o.debug == 1) {  // stereo_vision.cu:102-107 replaces XR / XT by 3 numbers each; only XT survives as a vector
        fprintf(stderr, "-d 1 (debug camera-to-robot transform) is not supported; using the calibration file's XR / XT\n");
    }
    svb_set_calibration(ctx, Q, cal.XR, cal.XT);
    printf("CUDA Init done\n");
    // extrapolate_point_cloud (stereo_vision.cu:523-524,245-265): the u8 map and the left image are resized by the factor and
    // the larger map is projected with the SAME Q
    const int PW = W * o.extrapolate, PH = H * o.extrapolate;
    const size_t PN = (size_t)PW * PH;
    double *points = (double *)svb_host_alloc(PN * 24);
    if (!points) {
        fprintf(stderr, "%s\n", svb_last_error());
        return 1;
    }
    std::vector<uint8_t> left, right, left_in, right_in, dmap(N), dmap_big, color_big;
    if (o.extrapolate != 1) {
        dmap_big.resize(PN);
        color_big.resize(PN * 4);
    }
    const bool resize = W != inW || H != inH;
    double FPS = 0;
    // imageLoop reads and decodes frame i+1 while frame i is on the GPU (the reference's loop is synchronous, stereo_vision.cu:645-697;
    // the per-frame timer starts behind imread there as well, so the printed FPS figures mean the same thing)
    struct DecodedPair {
        std::vector<uint8_t> l, r;
        bool ok = false;
    };
    auto decode_pair = [&o, inW, inH](unsigned i) {
        DecodedPair d;
        d.ok = load_bgra(frame_path(o.kitti_path, "image_02", i), inW, inH, &d.l) && load_bgra(frame_path(o.kitti_path, "image_03", i), inW, inH, &d.r);
        return d;
    };
    std::future<DecodedPair> ahead;
    if (max_files > 0) ahead = std::async(std::launch::async, decode_pair, 0u);
    for (unsigned i = 0; i < max_files; i++) {
        DecodedPair cur = ahead.get();
        if (i + 1 < max_files) ahead = std::async(std::launch::async, decode_pair, i + 1);
        if (!cur.ok) {
            if (ahead.valid()) ahead.wait();
            break;
        }
        (resize ? left_in : left).swap(cur.l);
        (resize ? right_in : right).swap(cur.r);
        const auto t0 = std::chrono::steady_clock::now();  // start_timer(t_start) after imread (:664)
        if (resize) {  // resize(left_img, left_img_OLD, out_img_size) (:665,676)
            left.resize(N * 4);
            right.resize(N * 4);
            if (svb_resize_bgra(left_in.data(), inW, inH, left.data(), W, H) != SVB_OK ||
                svb_resize_bgra(right_in.data(), inW, inH, right.data(), W, H) != SVB_OK) {
                fprintf(stderr, "%s\n", svb_last_error());
                return 1;
            }
        }
        double times[2] = {0, 0};
        const int rc = svb_point_cloud_bgra(ctx, left.data(), right.data(), points, dmap.data(), nullptr, times);
        if (rc == SVB_ERR_FEW_SUPPORT)
            printf("ERROR: Need at least 3 support points!\n");
        else if (rc != SVB_OK)
            fprintf(stderr, "%s\n", svb_last_error());
        if (o.extrapolate != 1 && rc == SVB_OK) {
            const auto p0 = std::chrono::steady_clock::now();
            if (svb_resize_bgra(left.data(), W, H, color_big.data(), PW, PH) != SVB_OK ||
                svb_resize_gray(dmap.data(), W, H, dmap_big.data(), PW, PH) != SVB_OK ||
                svb_reproject_u8(dmap_big.data(), PW, PH, Q, cal.XR, cal.XT, points) != SVB_OK)
                fprintf(stderr, "%s\n", svb_last_error());
            times[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - p0).count();
        }
        const double t_t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("(FPS=%f) (%d, %d) (t_t=%f, dmap_t=%f, pc_t=%f)\n", 1 / t_t, H, W, t_t, times[0] * 1e-3, times[1] * 1e-3);
        FPS += 1 / t_t;
        if (!o.dump_dir.empty()) {
            char name[64];
            snprintf(name, sizeof(name), "/%010u_disp.pgm", i);
            std::string err;
            if (!svb::write_pgm(o.dump_dir + name, dmap.data(), W, H, &err)) fprintf(stderr, "%s\n", err.c_str());
            if (o.extrapolate != 1) {  // the resized map and the cloud projected from it (raw float64 x, y, z)
                snprintf(name, sizeof(name), "/%010u_disp_e.pgm", i);
                if (!svb::write_pgm(o.dump_dir + name, dmap_big.data(), PW, PH, &err)) fprintf(stderr, "%s\n", err.c_str());
                snprintf(name, sizeof(name), "/%010u_points_e.f64", i);
                if (FILE *fp = fopen((o.dump_dir + name).c_str(), "wb")) {
                    fwrite(points, 24, PN, fp);
                    fclose(fp);
                }
            }
        }
    }
    printf("AVG_FPS=%f\n", max_files ? FPS / max_files : 0.0);
    svb_host_free(points);
    svb_destroy(ctx);
    clean();  // prints "Program exitted successfully!" and exits 0 (stereo_vision.cu:114-126)
    return 0;
}
