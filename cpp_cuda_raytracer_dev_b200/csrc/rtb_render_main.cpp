// rtb_render_main.cpp -- headless Linux driver: the reference's WinMain.cpp:69-237 without the window.
//
//   rtb_render --mesh FILE|geodesic:NU [--mode 0|1|2|-1] [--res WxH] [--frames N] [--zoom K]
//              [--rotate X,Y,Z,W] [--out PREFIX] [--png] [--cam px,py,pz,lx,ly,lz,ux,uy,uz] [--sweep]
//              [--objects N] [--shadows] [--samples n] [--light x,y,z]...   (scene extension: Camera::render())
//
// Same call sequence as the reference app: Camera(...) -> read_ply -> colour table -> Trixel ->
// set_sorted_voxels -> create_kd -> Object -> add_object -> per frame { [transform]; render;
// color_pixels(PHONG); present; color_pixels(SET) }.  "Present" writes PREFIX_NNNN.ppm and
// PREFIX_NNNN.ids (raw int32 hit ids) instead of StretchDIBits.  --sweep renders all frames through the
// batched rtb_render_sweep call instead of the frame loop.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rtb_framework.hpp"

using namespace rtbfw;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    std::string mesh = "geodesic:64", out;
    int mode = 0, W = 960, H = 540, frames = 1, zoom = 0;
    bool sweep = false, png = false, shadows = false;
    int objects = 1, samples = 0;
    std::vector<float> lights;
    float quat[4] = {0.0f, 0.09950371902099893f, 0.0f, 0.9950371902099893f};  // WinMain.cpp:187 (R key)
    float cam[9] = {0.0f, 0.10f, -1.0f, 0.0f, 0.10f, 0.0f, 0.0f, 1.0f, 0.0f};  // WinMain.cpp:71-73
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--mesh") mesh = next();
        else if (a == "--mode") mode = std::atoi(next());
        else if (a == "--res") std::sscanf(next(), "%dx%d", &W, &H);
        else if (a == "--frames") frames = std::atoi(next());
        else if (a == "--zoom") zoom = std::atoi(next());
        else if (a == "--out") out = next();
        else if (a == "--sweep") sweep = true;
        else if (a == "--png") png = true;
        else if (a == "--shadows") shadows = true;
        else if (a == "--objects") objects = std::atoi(next());
        else if (a == "--samples") samples = std::atoi(next());
        else if (a == "--light") { float l[3] = {2, 2, 2}; std::sscanf(next(), "%f,%f,%f", &l[0], &l[1], &l[2]); lights.insert(lights.end(), l, l + 3); }
        else if (a == "--rotate") std::sscanf(next(), "%f,%f,%f,%f", &quat[0], &quat[1], &quat[2], &quat[3]);
        else if (a == "--cam") std::sscanf(next(), "%f,%f,%f,%f,%f,%f,%f,%f,%f", &cam[0], &cam[1], &cam[2], &cam[3], &cam[4], &cam[5], &cam[6], &cam[7], &cam[8]);
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (rtb_device_count() < 1) { std::fprintf(stderr, "rtb_render: no CUDA device (there is no CPU fallback)\n"); return 3; }

    const float ar = (float)W / (float)H;  // WinMain.cpp:29,67
    Camera* main_cam = new Camera(W, H, ar * 0.024f, 0.024f, 0.055f, cam[0], cam[1], cam[2], cam[3], cam[4], cam[5], cam[6], cam[7], cam[8]);

    double t0 = now_s();
    T_fp* points_for_trixels = nullptr;
    T_uint tot_num_trixels = 0;
    int rc;
    if (mesh.rfind("geodesic:", 0) == 0) {
        const float center[3] = {0.0f, 0.1f, 0.0f};
        rc = rtb_mesh_geodesic(std::atoi(mesh.c_str() + 9), 0.08f, center, 0.05f, 1234u, &points_for_trixels, &tot_num_trixels);
    } else {
        rc = read_ply(mesh.c_str(), &points_for_trixels, &tot_num_trixels, (u8)(mode < 0 ? 255 : mode));
    }
    if (rc != 0) { std::fprintf(stderr, "mesh: %s\n", rtb_last_error()); return 1; }
    Color color_for_trixels;  // WinMain.cpp:114-121
    std::vector<Color::radiance> rad(tot_num_trixels, Color::radiance{0.1f, 0.55f, 0.20f});
    color_for_trixels.rad = rad.data();
    std::printf("Time to Read Tree: %f seconds\n\t\t primitives: %u\n\n", now_s() - t0, tot_num_trixels);

    t0 = now_s();
    Trixel* trixel_list = new Trixel(tot_num_trixels, points_for_trixels, &color_for_trixels);
    trixel_list->set_sorted_voxels(nullptr, tot_num_trixels);
    if (trixel_list->create_kd() != 0) { std::fprintf(stderr, "create_kd: %s\n", rtb_last_error()); return 1; }
    std::printf("Total Time to build tree: %f seconds\n", now_s() - t0);
    Object* obj1 = new Object(trixel_list);
    if (main_cam->add_object(obj1) != 0) { std::fprintf(stderr, "add_object: %s\n", rtb_last_error()); return 1; }

    // scene extension: WinMain.cpp:153,156 registers a second object over the same Trixel (its render is commented out, :214);
    // here every further object is slid sideways by the Q key's translation (WinMain.cpp:198-201) so that it can be seen
    const bool scene_mode = objects > 1 || shadows || samples > 1 || !lights.empty();
    std::vector<Object*> more;
    Input input;
    for (int k = 1; k < objects; k++) {
        Object* o = new Object(trixel_list);
        if (main_cam->add_object(o) != 0) { std::fprintf(stderr, "add_object: %s\n", rtb_last_error()); return 1; }
        for (int step = 0; step < 12 * k; step++) {
            input.set_quat(main_cam->o_prop.u.x, main_cam->o_prop.u.y, main_cam->o_prop.u.z, (k & 1) ? 0.012f : -0.012f);
            o->transform(&input, TRANSLATE_X);
        }
        more.push_back(o);
    }
    if (scene_mode) {
        if (!lights.empty() && main_cam->set_lights((int)(lights.size() / 3), lights.data()) != 0) { std::fprintf(stderr, "lights: %s\n", rtb_last_error()); return 1; }
        main_cam->set_shadows(shadows);
        main_cam->set_sample_rate(samples);
        if (sweep) { std::fprintf(stderr, "--sweep renders one object with the reference's shading; drop it for --objects/--shadows/--samples/--light\n"); return 2; }
    }
    for (int k = 0; k < zoom; k++) {  // W key, WinMain.cpp:190-193
        input.set_quat(main_cam->o_prop.n.x, main_cam->o_prop.n.y, main_cam->o_prop.n.z, 0.005f);
        obj1->transform(&input, TRANSLATE_Z);
    }
    auto present = [&](int frame, const uint32_t* c, const int32_t* id) {
        if (out.empty()) return;
        char name[64];
        std::snprintf(name, sizeof name, "_%04d", frame);
        write_ppm(out + name + (png ? ".png" : ".ppm"), c, (uint32_t)W, (uint32_t)H);
        write_raw_ids(out + name + ".ids", id, (uint64_t)W * H);
    };

    t0 = now_s();
    if (sweep) {
        std::vector<float> ops(5 * (size_t)frames);
        for (int f = 0; f < frames; f++) { ops[5 * f] = f ? (float)ROTATE_TRI_PY : 0.0f; std::memcpy(&ops[5 * f + 1], quat, sizeof quat); }
        // pinned buffers: the sweep then pre-fills the background on the host and the kernel stores the rest (no copy engine)
        const size_t words = (size_t)frames * W * H;
        void *colors_mem = nullptr, *ids_mem = nullptr;
        if (rtb_host_alloc(4 * words, &colors_mem) != 0 || rtb_host_alloc(4 * words, &ids_mem) != 0) {
            std::fprintf(stderr, "host_alloc: %s\n", rtb_last_error());
            return 1;
        }
        uint32_t* colors = static_cast<uint32_t*>(colors_mem);
        int32_t* ids = static_cast<int32_t*>(ids_mem);
        t0 = now_s();
        if (rtb_render_sweep(obj1->handle, main_cam->handle, frames, 1, ops.data(), RTB_RENDER_DEFAULT, colors, ids) != 0) {
            std::fprintf(stderr, "render_sweep: %s\n", rtb_last_error());
            return 1;
        }
        const double dt = now_s() - t0;
        for (int f = 0; f < frames; f++) present(f, colors + (size_t)f * W * H, ids + (size_t)f * W * H);
        rtb_host_free(colors_mem); rtb_host_free(ids_mem);
        std::printf("Resolution: %d x %d\nFPS: %f (sweep of %d frames, %.1f Mrays/s)\n", W, H, frames / dt, frames, frames * (double)W * H / dt / 1e6);
    } else {
        main_cam->color_pixels(SET_COLOR_TAG);
        for (int f = 0; f < frames; f++) {
            if (f) {  // R key held down, WinMain.cpp:186-189
                input.set_quat(quat[0], quat[1], quat[2], quat[3]);
                obj1->transform(&input, ROTATE_TRI_PY);
            }
            if (scene_mode) main_cam->render();       // Camera::render(): all objects, lights, shadows, samples
            else obj1->render(main_cam);              // WinMain.cpp:212
            main_cam->color_pixels(PHONG_COLOR_TAG);  // WinMain.cpp:213
            present(f, main_cam->h_color.c, main_cam->h_color.id);
            main_cam->color_pixels(SET_COLOR_TAG);    // WinMain.cpp:237
        }
        const double dt = now_s() - t0;
        std::printf("Resolution: %d x %d\nFPS: %f (frame loop of %d frames incl. PPM output)\n", W, H, frames / dt, frames);
    }
    for (Object* o : more) delete o;
    delete obj1; delete trixel_list; delete main_cam;
    rtb_free(points_for_trixels);
    return 0;
}
