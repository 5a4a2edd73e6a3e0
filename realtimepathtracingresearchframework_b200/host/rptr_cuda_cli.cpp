// rptr_cuda_cli -- headless driver of the CUDA backend over its C ABI (include/rptr_cuda.h): replays what `rptr --backend cuda
// --disable-ui --validation <prefix> --validation-spp N --pfm [--profiling <name>]` does once the two registration lines of
// INTEGRATION.md are in place, without the GLFW window the reference's main() always opens (main.cpp:66-202):
//   * the frame loop of run_app in validation mode (app.cpp:334-484; libapp/app_state.h:90-99): frames of
//     next_frame_spp(batch_spp) samples, reset_accumulation on the first, until accumulated_spp reaches the target;
//   * handle_mode_actions (libapp/app_state.cpp:464-481): "<prefix>_<%04d accumulated_spp>.pfm" through write_pfm
//     (util/write_image.cpp:34-66) when the frame is ready;
//   * BenchmarkInfo (libapp/benchmark_info.cpp:69-124): "<name>.csv" with the columns frames_total, keyframe,
//     frames_accumulated, render_time_ms, app_time_ms, one row per frame (RenderStats::render_time -> render_time_ms).
// Scenes are procedural (the .vks loader stays on the reference's side): BASELINE.json's Cornell box and random-triangle soup,
// generated exactly like realtimepathtracingresearchframework_b200/scenes.py (same splitmix64 stream, same 21-bit grid), so
// the images are bit-identical to those of the Python harness and of the oracle.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rptr_cuda.h"
#include "sky_fits.inc"
#include "vks_loader.hpp"

namespace {

struct GeometryData { std::vector<uint64_t> qverts; float scale[3], offset[3]; };
struct SceneData {
    std::vector<GeometryData> geometries;
    std::vector<rptr_base_material> materials;
    rptr_camera_params camera;
};

uint64_t pack_qvert(int64_t x, int64_t y, int64_t z) { // librender/quantize.h:7-11
    return ((uint64_t)x & 0x1FFFFFull) | (((uint64_t)y & 0x1FFFFFull) << 21) | (((uint64_t)z & 0x1FFFFFull) << 42);
}
int64_t snap(double v, double base, double scale) {
    const double g = std::floor((v - base) / scale);
    return (int64_t)(g < 0.0 ? 0.0 : (g > 2097151.0 ? 2097151.0 : g));
}
rptr_base_material material(float r, float g, float b) { // BaseMaterial defaults of types.py
    rptr_base_material m;
    memset(&m, 0, sizeof(m));
    m.base_color[0] = r; m.base_color[1] = g; m.base_color[2] = b;
    m.normal_map = -1;
    m.roughness = 1.0f; m.specular = 0.5f; m.clearcoat_gloss = 0.1f; m.ior = 1.5f;
    m.transmission_color[0] = m.transmission_color[1] = m.transmission_color[2] = 1.0f;
    return m;
}
rptr_camera_params look_at(const float *eye, const float *center, float fovy) { // scenes.look_at_camera
    rptr_camera_params c;
    float d[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
    const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); // numpy: float32 dot, float32 sqrt
    for (int k = 0; k < 3; ++k) { c.pos[k] = eye[k]; c.dir[k] = d[k] / len; c.up[k] = k == 1 ? 1.0f : 0.0f; }
    c.fovy = fovy;
    return c;
}

// SURVEY 8d "C1 synthetic input": scenes.cornell_box()
SceneData cornell_box() {
    SceneData s;
    const double base[3] = {-1.0, 0.0, -1.0}, scale = std::ldexp(1.0, -20);
    typedef double P[3];
    auto quad = [&](GeometryData &g, const P &a, const P &b, const P &c, const P &d) {
        const P *order[6] = {&a, &b, &c, &a, &c, &d};
        for (const P *p : order) g.qverts.push_back(pack_qvert(snap((*p)[0], base[0], scale), snap((*p)[1], base[1], scale), snap((*p)[2], base[2], scale)));
    };
    s.geometries.resize(4);
    for (GeometryData &g : s.geometries)
        for (int k = 0; k < 3; ++k) { g.scale[k] = (float)scale; g.offset[k] = (float)(base[k] + std::ldexp(1.0, -21)); }
    const P f0 = {-1, 0, -1}, f1 = {-1, 0, 1}, f2 = {1, 0, 1}, f3 = {1, 0, -1};
    const P c0 = {-1, 2, -1}, c1 = {1, 2, -1}, c2 = {1, 2, 1}, c3 = {-1, 2, 1};
    quad(s.geometries[0], f0, f1, f2, f3);                                   // floor
    quad(s.geometries[0], c0, c1, c2, c3);                                   // ceiling
    quad(s.geometries[0], f0, f3, c1, c0);                                   // back
    quad(s.geometries[1], f0, c0, c3, f1);                                   // left
    quad(s.geometries[2], f3, f2, c2, c1);                                   // right
    const P l0 = {-0.25, 1.998, -0.25}, l1 = {0.25, 1.998, -0.25}, l2 = {0.25, 1.998, 0.25}, l3 = {-0.25, 1.998, 0.25};
    quad(s.geometries[3], l0, l1, l2, l3);                                   // light, facing down
    const float cols[4][3] = {{0.73f, 0.73f, 0.73f}, {0.63f, 0.065f, 0.05f}, {0.14f, 0.45f, 0.091f}, {1.0f, 0.85f, 0.6f}};
    for (int j = 0; j < 4; ++j) {
        rptr_base_material m = material(cols[j][0], cols[j][1], cols[j][2]);
        m.ior = 1.0f; m.roughness = 1.0f; m.metallic = 0.0f; m.flags = RPTR_BASE_MATERIAL_NOALPHA;
        if (j == 3) m.emission_intensity = 17.0f;
        s.materials.push_back(m);
    }
    const float eye[3] = {0.0f, 1.0f, 3.4f}, center[3] = {0.0f, 1.0f, 0.0f};
    s.camera = look_at(eye, center, 40.0f);
    return s;
}

// SURVEY 8d "C2/C3/C5 input": scenes.random_triangles(n) -- splitmix64(seed 0x5EED1A7B200), 16 geometries, diffuse + GGX
SceneData random_triangles(int64_t n_tris) {
    SceneData s;
    const double box = 10.0, edge = 0.15, scale = std::ldexp(1.0, -16), base = -16.0;
    const uint64_t seed = 0x5EED1A7B200ull;
    auto uniform = [&](uint64_t i) { // scenes.splitmix64_uniform: value i (0-based) of the stream
        uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    };
    const int n_geometries = 16;
    const int64_t per = (n_tris + n_geometries - 1) / n_geometries;
    static const float palette[16][3] = {{0.80f, 0.80f, 0.80f}, {0.90f, 0.35f, 0.25f}, {0.25f, 0.65f, 0.90f}, {0.95f, 0.85f, 0.35f}, {0.35f, 0.80f, 0.45f},
                                         {0.75f, 0.40f, 0.85f}, {0.95f, 0.60f, 0.20f}, {0.30f, 0.35f, 0.85f}, {0.60f, 0.60f, 0.60f}, {0.85f, 0.25f, 0.45f},
                                         {0.20f, 0.75f, 0.75f}, {0.70f, 0.75f, 0.30f}, {0.90f, 0.90f, 0.95f}, {0.55f, 0.35f, 0.25f}, {0.40f, 0.55f, 0.35f},
                                         {0.50f, 0.50f, 0.70f}};
    for (int j = 0; j < n_geometries; ++j) {
        const int64_t lo = j * per, hi = std::min<int64_t>(n_tris, (j + 1) * per);
        if (lo >= hi) break;
        GeometryData g;
        for (int k = 0; k < 3; ++k) { g.scale[k] = (float)scale; g.offset[k] = (float)(base + std::ldexp(1.0, -17)); }
        for (int64_t t = lo; t < hi; ++t) {
            double u[9];
            for (int k = 0; k < 9; ++k) u[k] = uniform((uint64_t)t * 9 + k);
            double c[3], e1[3], e2[3];
            for (int k = 0; k < 3; ++k) { c[k] = (u[k] * 2.0 - 1.0) * box; e1[k] = (u[3 + k] * 2.0 - 1.0) * edge; e2[k] = (u[6 + k] * 2.0 - 1.0) * edge; }
            g.qverts.push_back(pack_qvert(snap(c[0], base, scale), snap(c[1], base, scale), snap(c[2], base, scale)));
            g.qverts.push_back(pack_qvert(snap(c[0] + e1[0], base, scale), snap(c[1] + e1[1], base, scale), snap(c[2] + e1[2], base, scale)));
            g.qverts.push_back(pack_qvert(snap(c[0] + e2[0], base, scale), snap(c[1] + e2[1], base, scale), snap(c[2] + e2[2], base, scale)));
        }
        s.geometries.push_back(std::move(g));
        rptr_base_material m = material(palette[j % 16][0], palette[j % 16][1], palette[j % 16][2]);
        m.roughness = (float)(0.1 + 0.05 * (j % 16)); m.metallic = (float)(j & 1); m.ior = 1.5f; m.specular = 0.5f; m.flags = RPTR_BASE_MATERIAL_NOALPHA;
        s.materials.push_back(m);
    }
    const float eye[3] = {0.0f, 0.0f, 30.0f}, center[3] = {0.0f, 0.0f, 0.0f};
    s.camera = look_at(eye, center, 65.0f);
    return s;
}

const char *USAGE =
    "usage: rptr_cuda_cli [--scene cornell|random:<triangles>|<file>.vks] [--img <x> <y>] [--backend cuda] [--disable-ui]\n"
    "                     [--eye <x> <y> <z> --target <x> <y> <z> --fovy <deg>] [--transmission]\n"
    "                     --validation <prefix> [--validation-spp <n>] [--pfm] [--batch-spp <n>] [--profiling <name>]\n"
    "                     [--sky default|slanted] [--device <ordinal>] [--gpus <n>] [--bvh-builder 0|1]\n"
    "Renders time 0 of a procedural scene or of a .vks scene file (with its <name>_textures/ directory; the file has no camera:\n"
    "pass --eye) to <prefix>_<spp>.pfm like `rptr --validation` (libapp/app_state.cpp:464-481);\n"
    "--profiling writes the BenchmarkInfo columns of libapp/benchmark_info.cpp:69-124 to <name>.csv.  Needs a CUDA device:\n"
    "there is no CPU fallback.\n";

int die(const std::string &msg) {
    fprintf(stderr, "rptr_cuda_cli: %s\n", msg.c_str());
    return 1;
}

} // namespace

int main(int argc, char **argv) {
    std::string scene_name = "cornell", validation_prefix, profiling_name, sky = "default";
    bool scene_hash = false, have_eye = false, transmission = false;
    float eye[3] = {0.0f, 0.0f, 10.0f}, target[3] = {0.0f, 0.0f, 0.0f}, fovy = 65.0f;
    int width = 1920, height = 1080, target_spp = -1, batch_spp = 1, device = 0, gpus = 1, bvh_builder = 1;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char *what) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "rptr_cuda_cli: %s needs a value\n%s", what, USAGE); exit(2); }
            return argv[++i];
        };
        if (a == "--help" || a == "-h") { fputs(USAGE, stdout); return 0; }
        else if (a == "--scene") scene_name = next("--scene");
        else if (a == "--img") { width = atoi(next("--img")); height = atoi(next("--img")); }
        else if (a == "--backend") { if (std::string(next("--backend")) != "cuda") return die("only --backend cuda is served here"); }
        else if (a == "--disable-ui" || a == "--pfm") {}
        else if (a == "--exr") return die("EXR output needs the reference's tinyexr; use --pfm");
        else if (a == "--validation") validation_prefix = next("--validation");
        else if (a == "--validation-spp") { target_spp = atoi(next("--validation-spp")); if (target_spp < 1) target_spp = -1; } // cmdline.cpp:382-386
        else if (a == "--batch-spp") batch_spp = atoi(next("--batch-spp"));
        else if (a == "--profiling") profiling_name = next("--profiling");
        else if (a == "--sky") sky = next("--sky");
        else if (a == "--device") device = atoi(next("--device"));
        else if (a == "--gpus") gpus = atoi(next("--gpus"));
        else if (a == "--bvh-builder") bvh_builder = atoi(next("--bvh-builder"));
        else if (a == "--eye") { for (float &v : eye) v = (float)atof(next("--eye")); have_eye = true; }
        else if (a == "--target") { for (float &v : target) v = (float)atof(next("--target")); }
        else if (a == "--fovy") fovy = (float)atof(next("--fovy"));
        else if (a == "--transmission") transmission = true; // the GLTF_SUPPORT_TRANSMISSION build (option "transmission")
        else if (a == "--scene-hash") scene_hash = true; // FNV-1a of the generated scene (tests: equals scenes.py), no device needed
        else return die("unknown argument " + a + "\n" + USAGE);
    }
    if (validation_prefix.empty() && !scene_hash) return die(std::string("validation mode needs --validation <prefix>\n") + USAGE);
    if (target_spp < 0) target_spp = 1;
    if (batch_spp < 1 || gpus < 1 || gpus > 16 || width < 1 || height < 1) return die("invalid --batch-spp / --gpus / --img");
    const SkyFitEntry *fit = nullptr;
    for (const SkyFitEntry &e : k_sky_fits)
        if (sky == e.name) fit = &e;
    if (!fit) return die("unknown --sky " + sky);

    SceneData sd;
    rptr_host::VksScene vks;
    const bool from_file = scene_name.size() > 4 && (scene_name.rfind(".vks") == scene_name.size() - 4 || scene_name.rfind(".vkrs") == scene_name.size() - 5);
    if (from_file) { // librender/scene.cpp:61-62 dispatches on the extension
        try {
            rptr_host::load_vks(scene_name, vks);
        } catch (const std::exception &e) {
            return die(e.what());
        }
        if (scene_hash) { printf("%016llx\n", (unsigned long long)vks.hash()); return 0; }
        if (!have_eye) return die("a .vks file has no camera: pass --eye <x> <y> <z> [--target <x> <y> <z>] [--fovy <deg>]");
        sd.camera = look_at(eye, target, fovy);
    }
    else if (scene_name == "cornell") sd = cornell_box();
    else if (scene_name.rfind("random:", 0) == 0) sd = random_triangles(atoll(scene_name.c_str() + 7));
    else return die("unknown --scene " + scene_name);
    if (scene_hash) {
        uint64_t h = 1469598103934665603ull;
        auto eat = [&](const void *p, size_t n) {
            const unsigned char *b = (const unsigned char *)p;
            for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        };
        for (const GeometryData &g : sd.geometries) {
            eat(g.qverts.data(), g.qverts.size() * 8);
            eat(g.scale, 12);
            eat(g.offset, 12);
        }
        eat(sd.materials.data(), sd.materials.size() * sizeof(rptr_base_material));
        eat(&sd.camera, sizeof(sd.camera));
        printf("%016llx\n", (unsigned long long)h);
        return 0;
    }
    // the reference's in-memory scene model: one Mesh of all geometries, one ParameterizedMesh, one Instance (identity)
    std::vector<rptr_geometry_desc> geoms(sd.geometries.size());
    std::vector<int32_t> material_offsets(sd.geometries.size());
    for (size_t g = 0; g < geoms.size(); ++g) {
        memset(&geoms[g], 0, sizeof(geoms[g]));
        geoms[g].qverts = sd.geometries[g].qverts.data();
        geoms[g].n_tris = (int32_t)(sd.geometries[g].qverts.size() / 3);
        for (int k = 0; k < 3; ++k) { geoms[g].quantized_scaling[k] = sd.geometries[g].scale[k]; geoms[g].quantized_offset[k] = sd.geometries[g].offset[k]; }
        material_offsets[g] = (int32_t)g;
    }
    rptr_mesh_desc mesh{0, (int32_t)geoms.size()};
    rptr_pmesh_desc pmesh{0, (int32_t)material_offsets.size(), material_offsets.data(), nullptr, 0};
    rptr_instance_desc inst;
    inst.pmesh_id = 0;
    const float identity[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    memcpy(inst.transform, identity, sizeof(identity));
    rptr_scene_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.geometries = geoms.data(); desc.n_geometries = (int32_t)geoms.size();
    desc.meshes = &mesh; desc.n_meshes = 1;
    desc.pmeshes = &pmesh; desc.n_pmeshes = 1;
    desc.instances = &inst; desc.n_instances = 1;
    desc.materials = sd.materials.data(); desc.n_materials = (int32_t)sd.materials.size();
    if (from_file) desc = vks.desc();
    else if (have_eye) sd.camera = look_at(eye, target, fovy);

    // one context per GPU; with several, a communicator over them (the reference's single render thread drives all devices)
    std::vector<rptr_ctx *> ctxs((size_t)gpus, nullptr);
    for (int g = 0; g < gpus; ++g)
        if (rptr_cuda_create(device + g, &ctxs[g]) != 0) return die(std::string("create: ") + rptr_cuda_last_error(nullptr));
#define CHECK(call, c)                                                                         \
    if ((call) != 0) return die(std::string(#call) + ": " + rptr_cuda_last_error(c))
    if (gpus > 1) CHECK(rptr_cuda_comm_init_all(ctxs.data(), gpus), ctxs[0]);
    rptr_light_sampling_config lighting{0.0f, 16, 15.0f, 0.0f}; // LightSamplingConfig defaults (librender/render_params.glsl.h:123-128)
    for (rptr_ctx *c : ctxs) {
        CHECK(rptr_cuda_initialize(c, width, height), c);
        CHECK(rptr_cuda_set_option(c, "bvh_builder", bvh_builder), c);
        if (transmission) CHECK(rptr_cuda_set_option(c, "transmission", 1), c);
        CHECK(rptr_cuda_set_scene(c, &desc, &lighting), c);
        CHECK(rptr_cuda_set_scene_params(c, &fit->params), c);
    }
    rptr_render_params params; // RenderParams defaults (librender/render_params.glsl.h:130-155)
    memset(&params, 0, sizeof(params));
    params.batch_spp = 1; params.max_path_depth = RPTR_MAX_PATH_DEPTH; params.rr_path_depth = RPTR_DEFAULT_RR_PATH_DEPTH;
    params.focus_distance = 2.5f; params.pixel_radius = 1.0f; params.variance_radius = 4.0f; params.early_tone_mapping_mode = -1;
    params.spp_accumulation_window = 8; params.render_upscale_factor = 1; params.focal_length = 35.0f;

    FILE *csv = nullptr;
    if (!profiling_name.empty()) {
        csv = fopen((profiling_name + ".csv").c_str(), "w");
        if (!csv) return die("cannot open " + profiling_name + ".csv");
        fprintf(csv, "frames_total,keyframe,frames_accumulated,render_time_ms,app_time_ms\n"); // BenchmarkInfo::open_csv
    }
    int accumulated_spp = 0, frames_total = 0;
    std::vector<float> pixels((size_t)width * height * 4);
    bool done_accumulating = false;
    while (!done_accumulating) {
        const auto t0 = std::chrono::steady_clock::now();
        // next_frame_spp (libapp/app_state.h:90-94)
        params.batch_spp = (target_spp > 0 && accumulated_spp > target_spp - batch_spp) ? target_spp - accumulated_spp : batch_spp;
        for (rptr_ctx *c : ctxs) {
            CHECK(rptr_cuda_begin_frame(c, &sd.camera, &params, &lighting, accumulated_spp == 0 ? 1 : 0, 0, 0.0), c); // app.cpp:360
            CHECK(rptr_cuda_draw_frame(c, 0), c);
            CHECK(rptr_cuda_end_frame(c, 0), c);
        }
        rptr_render_stats stats;
        CHECK(rptr_cuda_stats(ctxs[0], &stats), ctxs[0]);
        for (int g = 1; g < gpus; ++g) {
            rptr_render_stats other;
            CHECK(rptr_cuda_stats(ctxs[g], &other), ctxs[g]);
            if (other.render_time > stats.render_time) stats.render_time = other.render_time;
        }
        accumulated_spp = stats.spp; // update_accumulated_spp (libapp/app_state.h:95-99)
        done_accumulating = accumulated_spp >= target_spp;
        ++frames_total;
        const double app_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (csv) fprintf(csv, "%d,%d,%d,%g,%g\n", frames_total, 1, accumulated_spp, (double)stats.render_time, app_ms); // BenchmarkInfo::write_csv
    }
    if (csv) fclose(csv);
    // handle_mode_actions: <prefix>_<%04d accumulated_spp>, PFM written from readback_framebuffer(float*)
    if (gpus > 1) CHECK(rptr_cuda_reduce_framebuffer_all(ctxs.data(), gpus, 0), ctxs[0]);
    if (rptr_cuda_readback_f32(ctxs[0], pixels.size(), pixels.data()) != pixels.size()) return die(std::string("readback: ") + rptr_cuda_last_error(ctxs[0]));
    char name[1024];
    snprintf(name, sizeof(name), "%s_%04d", validation_prefix.c_str(), accumulated_spp);
    if (rptr_write_pfm(name, (uint32_t)width, (uint32_t)height, 4, pixels.data()) != 0) return die(std::string("cannot write ") + name + ".pfm");
    printf("%s.pfm: %d x %d, %d spp in %d frames on %d GPU%s (%s)\n", name, width, height, accumulated_spp, frames_total, gpus, gpus > 1 ? "s" : "", rptr_cuda_name());
    for (rptr_ctx *c : ctxs) rptr_cuda_destroy(c);
    return 0;
}
