/* c_abi_demo.c -- the C ABI of include/rrt_b200.h driven from plain C (no Python, no torch):
 * renders match_mirror.py's scene (two spheres + a square, Phong, 128x128, 4 AA samples,
 * in-kernel jitter) with rrt_render_forward, runs the fused forward + MSE + reverse pass against
 * the left-right flipped image (match_mirror.py:40,45) with rrt_render_fused_mse, and writes
 * frame0.ppm.  Build (see tests/test_gpu_api.py::test_c_abi_demo):
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/c_abi_demo.c -o c_abi_demo \
 *       -Lreversible_raytracer_b200 -lrrt_b200 -L/usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/reversible_raytracer_b200
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rrt_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define RK(x) do { int r_ = (x); if (r_ != RRT_OK) { fprintf(stderr, "%s: %s\n", #x, rrt_last_error()); return 3; } } while (0)

static void* upload(const void* src, size_t bytes) {
    void* d = NULL;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return NULL;
    if (cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
    return d;
}

int main(int argc, char** argv) {
    const int n = 128, S = 4, N = 3;
    /* w2o rows 0..2 of each shape (shape.py:74-75): translate(c).inverse() = translate(-c);
     * the square is (translate((0,0,3)) * rotate(50, y)).inverse() = rotate^T * translate(-c) */
    const double a = 50.0 * M_PI / 180.0, cs = cos(a), sn = sin(a);
    float w2o[3][12] = {
        {1, 0, 0, 0.5f, 0, 1, 0, 0.5f, 0, 0, 1, -4.f},
        {1, 0, 0, -0.5f, 0, 1, 0, -0.5f, 0, 0, 1, -4.f},
        {0}};
    /* R = rotate(50, (0,1,0)): [[c,0,s],[0,1,0],[-s,0,c]] (transform.py:95-122); w2o = [R^T | -R^T c] */
    float Rt[9] = {(float)cs, 0, (float)-sn, 0, 1, 0, (float)sn, 0, (float)cs};
    const float c[3] = {0, 0, 3};
    for (int r = 0; r < 3; r++) {
        for (int q = 0; q < 3; q++) w2o[2][r * 4 + q] = Rt[r * 3 + q];
        w2o[2][r * 4 + 3] = -(Rt[r * 3] * c[0] + Rt[r * 3 + 1] * c[1] + Rt[r * 3 + 2] * c[2]);
    }
    /* material rows: ka, kd, ks, shininess, r, g, b (scene.py:89-101) */
    const float mat[3][7] = {{0.5f, 0.7f, 0.3f, 50.f, 0.2f, 0.9f, 0.4f},
                             {0.4f, 0.9f, 0.3f, 50.f, 0.87f, 0.1f, 0.507f},
                             {0.4f, 0.9f, 0.3f, 50.f, 0.87f, 0.1f, 0.507f}};
    const float light[6] = {-1.f, -1.f, 2.f, 1.f, 0.87f, 0.961f};
    const float camera[15] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 1};   /* identity o2w + look_at */
    const int32_t types[3] = {RRT_OBJ_SPHERE, RRT_OBJ_SPHERE, RRT_OBJ_SQUARE};

    rrt_scene sc;
    memset(&sc, 0, sizeof sc);
    sc.n = n; sc.samples = S; sc.num_objects = N; sc.num_scenes = 1;
    sc.shader = RRT_SHADER_PHONG; sc.transpose = 1; sc.max_depth = 1.f; sc.seed = 3;
    sc.obj_type = (const int32_t*)upload(types, sizeof types);
    sc.w2o = (const float*)upload(w2o, sizeof w2o);
    sc.material = (const float*)upload(mat, sizeof mat);
    sc.light = (const float*)upload(light, sizeof light);
    sc.camera = (const float*)upload(camera, sizeof camera);
    if (!sc.obj_type || !sc.w2o || !sc.material || !sc.light || !sc.camera) { fprintf(stderr, "upload failed\n"); return 2; }

    const size_t px = (size_t)n * n * 3, rays = (size_t)S * n * n;
    float *d_image, *d_target, *d_grad;
    int32_t* d_hit;
    double* d_loss;
    CK(cudaMalloc((void**)&d_image, px * sizeof(float)));
    CK(cudaMalloc((void**)&d_target, px * sizeof(float)));
    CK(cudaMalloc((void**)&d_hit, rays * sizeof(int32_t)));
    CK(cudaMalloc((void**)&d_grad, RRT_GRAD_SIZE(N) * sizeof(float)));
    CK(cudaMalloc((void**)&d_loss, sizeof(double)));

    RK(rrt_render_forward(&sc, d_image, d_hit, NULL, NULL));
    float* image = (float*)malloc(px * sizeof(float));
    float* flipped = (float*)malloc(px * sizeof(float));
    int32_t* hit = (int32_t*)malloc(rays * sizeof(int32_t));
    CK(cudaMemcpy(image, d_image, px * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hit, d_hit, rays * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
            memcpy(flipped + ((size_t)y * n + x) * 3, image + ((size_t)y * n + (n - 1 - x)) * 3, 3 * sizeof(float));
    CK(cudaMemcpy(d_target, flipped, px * sizeof(float), cudaMemcpyHostToDevice));

    RK(rrt_render_fused_mse(&sc, d_target, NULL, NULL, NULL, d_loss, d_grad, NULL));
    double loss = 0;
    float grad[3 * RRT_OBJ_GRAD_STRIDE + RRT_GLOBAL_GRAD];
    CK(cudaMemcpy(&loss, d_loss, sizeof loss, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(grad, d_grad, sizeof grad, cudaMemcpyDeviceToHost));

    size_t hits[4] = {0, 0, 0, 0};
    for (size_t r = 0; r < rays; r++) hits[hit[r] < 0 ? 3 : hit[r]]++;
    printf("version %d  loss %.6f  hits sphere0 %zu sphere1 %zu square %zu background %zu\n", rrt_version(), loss,
           hits[0], hits[1], hits[2], hits[3]);
    /* d loss / d centre = -(d loss / d w2o[:,3]) for a pure translation */
    printf("dloss/dcentre0 = (%.5f, %.5f, %.5f)  dloss/dcentre1 = (%.5f, %.5f, %.5f)\n", -grad[3], -grad[7], -grad[11],
           -grad[RRT_OBJ_GRAD_STRIDE + 3], -grad[RRT_OBJ_GRAD_STRIDE + 7], -grad[RRT_OBJ_GRAD_STRIDE + 11]);
    const char* out = argc > 1 ? argv[1] : "frame0.ppm";
    FILE* f = fopen(out, "wb");
    if (f) {
        fprintf(f, "P6\n%d %d\n255\n", n, n);
        for (size_t q = 0; q < px; q++) {
            float v = image[q] < 0 ? 0 : (image[q] > 1 ? 1 : image[q]);
            fputc((int)(v * 255.0f + 0.5f), f);
        }
        fclose(f);
    }
    return 0;
}
