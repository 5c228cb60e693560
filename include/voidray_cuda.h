/* voidray_cuda.h — C ABI of the B200-native replacement for voidray's CPU render loop.
 *
 * The reference (LevKruglyak/voidray) has no FFI: its "operator API" is a set of Rust functions and
 * traits. Each entry point below names the reference seam it replaces (paths relative to the
 * reference checkout). INTEGRATION.md shows the Rust `extern "C"` block + safe wrapper a maintainer
 * would add to voidray_renderer to route `iterative_render` through this library.
 *
 * Conventions
 *  - every function returns 0 (VR_OK) or a negative vr_status; vr_last_error() gives the message of
 *    the calling thread's last failure. Nothing panics, aborts or throws across the boundary.
 *  - all pointers are borrowed for the duration of the call and copied; the library owns device
 *    memory, the caller owns host memory. Handles are uint32_t indices exactly like the reference's
 *    `usize` newtypes (core/scene.rs:46-59).
 *  - all arithmetic is f32 (`Float = f32`, util/vector.rs:11-16). There is no CPU fallback: every
 *    call that needs the device fails with VR_ERR_CUDA when no sm_100 GPU is usable.
 *  - threading: one render thread calls vr_render_accumulate; vr_render_cancel / vr_render_stats may
 *    be called from any thread; vr_render_read_accum / vr_render_resolve must not overlap an
 *    accumulate on the same vr_render (mirrors the target write-lock, render/target.rs:266-268).
 */
#ifndef VOIDRAY_CUDA_H
#define VOIDRAY_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vr_context vr_context;
typedef struct vr_scene vr_scene;
typedef struct vr_render vr_render;

typedef enum vr_status {
    VR_OK = 0,
    VR_ERR_INVALID = -1, /* bad argument / handle / call order */
    VR_ERR_CUDA = -2,    /* CUDA runtime failure, or no usable device */
    VR_ERR_OOM = -3,
    VR_ERR_CANCELLED = -4 /* accumulate stopped early by vr_render_cancel */
} vr_status;

const char* vr_last_error(void);
/* ABI version, bumped on any signature/layout change. */
uint32_t vr_abi_version(void);

/* ---- context: one CUDA device + one stream ------------------------------------------------- */
/* `cuda_stream` is a cudaStream_t (or NULL: the library creates its own non-blocking stream). All
 * kernels and copies of this context are issued on that stream, so a caller that passes its own
 * stream can bracket calls with its own CUDA events. */
int32_t vr_context_create(int32_t device, void* cuda_stream, vr_context** out);
int32_t vr_context_destroy(vr_context* ctx);
/* A device group in ONE process (the Rust host of the reference is one process with one render thread,
 * render/renderer.rs:35-125): `device_ids[0..n_devices)` are CUDA device ordinals, each with a library-owned
 * stream. The returned context stands for the whole group and is used like any other:
 *   vr_scene_commit        flattens once and uploads the scene to every device (one host thread per device);
 *   vr_render_accumulate   cuts the call's sample range into n_devices contiguous pieces, one per device (the
 *                          random streams are keyed by (pixel, global sample index), so the union is the sample
 *                          set one device would draw); blocking until every device is done;
 *   vr_render_read_accum / vr_render_resolve
 *                          sum the devices' accumulation buffers on device_ids[0] — the other devices' buffers
 *                          are read over peer memory (cudaDeviceEnablePeerAccess: NVLink / NVSwitch) inside the
 *                          reduce(+tonemap) kernel, in device order; non-destructive, so progressive display works;
 *   the vr_debug_* gates, vr_render_accum_device_ptr and the IPC reduce below see device_ids[0] only.
 * The result equals a one-device render up to f32 summation order. A device may be listed more than once (several
 * shards on one GPU; of use for testing on a one-GPU machine). At most 16 entries. */
int32_t vr_context_create_multi(const int32_t* device_ids, uint32_t n_devices, vr_context** out);
int32_t vr_context_device_count(vr_context* ctx, uint32_t* n_devices);

/* ---- scene builder: core/scene.rs:94-161 --------------------------------------------------- */
/* Scene::empty() (scene.rs:95-111): default camera look_at((1,0,10),(0,0,0),(0,1,0), PI/6), no
 * environment. */
int32_t vr_scene_create(vr_context* ctx, vr_scene** out);
int32_t vr_scene_destroy(vr_scene* scene);

/* Scene::add_image_texture (scene.rs:154-160) after `image::open(path).to_rgb32f()`
 * (core/texture.rs:36-49): w*h RGB f32 texels, row 0 first. sample_type: 0 Nearest, 1 Bilinear
 * (texture.rs:23-26). */
int32_t vr_scene_add_texture_rgb32f(vr_scene* scene, const float* rgb, uint32_t w, uint32_t h,
                                    int32_t sample_type, uint32_t* texture);

/* The same for decoded 8- / 16-bit images: `to_rgb32f` is channel / 255 resp. / 65535, alpha dropped
 * (image 0.24.3). `channels` is 3 or 4 (RGBA: the fourth channel is skipped). */
int32_t vr_scene_add_texture_rgb8(vr_scene* scene, const uint8_t* pixels, uint32_t w, uint32_t h, uint32_t channels,
                                  int32_t sample_type, uint32_t* texture);
int32_t vr_scene_add_texture_rgb16(vr_scene* scene, const uint16_t* pixels, uint32_t w, uint32_t h, uint32_t channels,
                                   int32_t sample_type, uint32_t* texture);

/* Scene::add_image_texture(path, sample_type) (scene.rs:154-160 -> ImageTexture::new, core/texture.rs:36-49):
 * decodes the file in the library like `image::open(path).unwrap().to_rgb32f()` — PNG, baseline / progressive
 * JPEG, TIFF (strips and tiles; none / LZW / Deflate / PackBits), BMP, GIF, ICO, DDS (DXT1/3/5), TGA, PNM, farbfeld, Radiance HDR, scan-line / tiled OpenEXR
 * (none / RLE / ZIPS / ZIP / PIZ / PXR24 / B44 / B44A), recognised by magic bytes (TGA by its extension). Where the
 * reference panics (missing or undecodable file) this returns VR_ERR_INVALID with the reason in vr_last_error(). */
int32_t vr_scene_add_image_texture_file(vr_scene* scene, const char* path, int32_t sample_type, uint32_t* texture);

/* Host-only: the decoder behind the *_file calls. *rgb receives a malloc'ed w*h*3 f32 array (row 0 first), to be
 * released with vr_image_free. Needs no CUDA device. */
int32_t vr_image_load_rgb32f(const char* path, uint32_t* w, uint32_t* h, float** rgb);
int32_t vr_image_free(float* rgb);

/* Scene::add_mesh(Arc::new(Mesh::from_buffers(vertices, indices))) (scene.rs:128-135,
 * core/mesh.rs:76-116). positions 3*n_vertices, uvs 2*n_vertices, normals 3*n_vertices (uvs /
 * normals may be NULL = zeros, like Vertex::position, mesh.rs:26-32); indices: n_indices u32,
 * consumed in chunks of 3 (a trailing partial chunk is ignored like chunks_exact). */
int32_t vr_scene_add_mesh(vr_scene* scene, const float* positions, const float* uvs, const float* normals,
                          uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices,
                          uint32_t* surface);
/* Scene::add_mesh_from_file(path) (scene.rs:137-144 -> Mesh::from_file, core/mesh.rs:46-74): Wavefront OBJ with
 * obj-rs 0.7.0 `load_obj::<TexturedVertex, u32>` semantics — every face must be a triangle of v/vt/vn triples,
 * each distinct triple becomes one vertex in first-seen order, (u, v) = the first two texture coordinates.
 * n_vertices / n_triangles (may be NULL) receive what the reference prints (mesh.rs:67-72). */
int32_t vr_scene_add_mesh_from_obj_file(vr_scene* scene, const char* path, uint32_t* surface, uint32_t* n_vertices,
                                        uint32_t* n_triangles);

/* Host-only: the loader behind vr_scene_add_mesh_from_obj_file (needs no CUDA device). The arrays are malloc'ed:
 * positions 3*n_vertices, uvs 2*n_vertices, normals 3*n_vertices, indices n_indices; release with vr_obj_free. */
typedef struct vr_obj_mesh {
    uint32_t n_vertices, n_indices;
    float* positions;
    float* uvs;
    float* normals;
    uint32_t* indices;
} vr_obj_mesh;
int32_t vr_obj_load(const char* path, vr_obj_mesh* out);
int32_t vr_obj_free(vr_obj_mesh* mesh);

/* Scene::add_analytic_surface(Surfaces::sphere(center, radius)) — voidray_common/src/surfaces.rs:18-20,31-80 */
int32_t vr_scene_add_sphere(vr_scene* scene, const float center[3], float radius, uint32_t* surface);
/* Scene::add_analytic_surface(Surfaces::ground_plane(height)) — surfaces.rs:22-24,82-114 */
int32_t vr_scene_add_ground_plane(vr_scene* scene, float height, uint32_t* surface);

/* The closed set of `dyn Material` implementations (voidray_common/src/simple.rs:17-58). */
typedef enum vr_material_kind {
    VR_MAT_LAMBERTIAN = 0,      /* Materials::lambertian / lambertian_texture(_no_normal), simple.rs:88-132 */
    VR_MAT_METAL = 1,           /* Materials::metal(albedo, fuzz), simple.rs:135-160; param = fuzz */
    VR_MAT_DIELECTRIC = 2,      /* Materials::dielectric(ir), simple.rs:187-231; param = ir */
    VR_MAT_EMISSION = 3,        /* Materials::colored_emissive(color, strength), simple.rs:163-184; param = strength */
    VR_MAT_LAMBERTIAN_BSDF = 4, /* Materials::lambertian_bsdf(albedo), simple.rs:60-81 via core/traits.rs:23-40 */
    VR_MAT_MICROFACET = 5       /* MicrofacetBSDF{color,index,roughness,metallic,emittance,transparent},
                                   voidray_common/src/microfacet.rs:9-313 via core/traits.rs:23-40 */
} vr_material_kind;

typedef struct vr_material_desc {
    int32_t kind;       /* vr_material_kind */
    float color[3];     /* albedo / emission colour */
    float param;        /* fuzz | ir | strength */
    int32_t albedo_tex; /* texture handle, or -1: use `color` (ColorType, simple.rs:83-86) */
    int32_t normal_tex; /* texture handle, or -1 (Lambertian.normal, simple.rs:90) */
    /* VR_MAT_MICROFACET only (microfacet.rs:9-27); `emittance` is carried but, as in the reference, unused by
       bsdf()/sample() */
    float index, roughness, metallic, emittance;
    int32_t transparent;
} vr_material_desc;

/* Scene::add_material (scene.rs:113-119) */
int32_t vr_scene_add_material(vr_scene* scene, const vr_material_desc* desc, uint32_t* material);
/* Scene::add_object(material, surface) (scene.rs:146-152). NOTE the reference resolves the material
 * of a hit through objects[surface_index] (scene.rs:183-184), not through the object's own surface
 * field; this library reproduces that. */
int32_t vr_scene_add_object(vr_scene* scene, uint32_t material, uint32_t surface, uint32_t* object);

/* `scene.camera = Camera{eye, direction, up, fov, dof}` (core/camera.rs:7-22). direction/up are used
 * as given (the mushroom example sets a non-unit direction). focal_point may be NULL if !has_dof. */
int32_t vr_scene_set_camera(vr_scene* scene, const float eye[3], const float direction[3], const float up[3],
                            float fov, int32_t has_dof, float aperture, const float focal_point[3]);
/* Camera::look_at(eye, center, up, fov) (camera.rs:26-36), dof = None */
int32_t vr_scene_set_camera_look_at(vr_scene* scene, const float eye[3], const float center[3],
                                    const float up[3], float fov);

/* The arithmetic of Camera::look_at alone (no scene needed): direction = normalize(center - eye),
 * up = normalize(up - up.dot(direction) * direction), in f32 with the reference's operation order. */
int32_t vr_camera_look_at(const float eye[3], const float center[3], const float up[3], float direction_out[3],
                          float up_out[3]);

/* scene.environment = Environments::uniform(rgb) — voidray_common/src/environments.rs:9-11,19-33 */
int32_t vr_scene_set_environment_uniform(vr_scene* scene, const float rgb[3]);
/* scene.environment = Environments::hdri(path) after image::open().to_rgb32f() — environments.rs:13-16,35-86 */
int32_t vr_scene_set_environment_hdri_rgb32f(vr_scene* scene, const float* rgb, uint32_t w, uint32_t h);
/* scene.environment = Environments::hdri(path) with the file decoded in the library (environments.rs:42-55);
 * formats as vr_scene_add_image_texture_file. */
int32_t vr_scene_set_environment_hdri_file(vr_scene* scene, const char* path);
/* scene.environment = None */
int32_t vr_scene_clear_environment(vr_scene* scene);

/* Accelerable::build_acceleration (scene.rs:163-179): flatten every mesh into the packed triangle /
 * BVH layout, build the BVH, upload scene to the device. Must be called before vr_render_begin and
 * again after any edit. */
int32_t vr_scene_commit(vr_scene* scene);

/* What the last vr_scene_commit built and uploaded (println! on load in the reference, core/mesh.rs:67). */
typedef struct vr_scene_info {
    uint32_t n_triangles;
    uint32_t n_bvh_nodes;
    uint32_t bvh_depth;
    uint32_t n_analytic_surfaces;
    uint32_t n_textures;
    uint32_t reserved;
    uint64_t h2d_bytes;   /* bytes copied host -> device by the commit (geometry, BVH, textures, environment) */
    double flatten_ms;    /* host: tie ranks + BVH build + packing */
    double upload_ms;     /* host -> device copies: what they add on top of the host build (the texture / environment
                             copies are issued first and overlap it) */
} vr_scene_info;
int32_t vr_scene_get_info(vr_scene* scene, vr_scene_info* out);

/* ---- render: render/iterative.rs:11-55, core/tracer.rs, core/settings.rs:15-33 ------------- */
typedef struct vr_render_settings {
    uint32_t total_samples; /* RenderSettings.total_samples: accumulated values are divided by this */
    uint32_t max_bounces;   /* RenderSettings.max_bounces */
    float firefly_clamp;    /* RenderSettings.firefly_clamp */
    int32_t render_mode;    /* 0 RenderMode::Full, 1 RenderMode::Normal (settings.rs:9-13) */
    int32_t pixel_mapping;  /* 0: y = index / width (intended); 1: the reference's y = index / height with
                               u32 wrapping (iterative.rs:26,33) — identical for square targets */
    int32_t integrator;     /* 0: the reference estimator (parity). 1: "fast" — the same integrand, but Lambertian
                               directions are drawn by one-sample MIS between the reference's normal + UnitSphere
                               density and an HDRI luminance table, and paths play Russian roulette from depth 3.
                               Not in the reference (core/tracer.rs:19-56 has neither); equal in expectation to
                               integrator 0 only without the firefly clamp. */
    uint64_t seed;          /* Philox4x32-10 key. The reference's thread_rng() is unseedable (iterative.rs:29) */
    uint32_t sample_offset; /* global index of this render's first camera sample: rank r of an N-GPU job
                               renders [sample_offset, sample_offset + its share) of every pixel */
    uint32_t max_paths_in_flight; /* paths in flight, shared by the two wavefronts consecutive batches alternate between;
                                     0 = library default (2 x 32 Mi paths, 248 B each at 8 bounces); < 64 Mi per wavefront.
                                     Diagnostic: the environment variable VOIDRAY_STREAMS=n (1..4, read by vr_render_begin)
                                     sets the number of wavefronts, VOIDRAY_CAMERA_CULL=0 / 1 pins the camera-ray culling
                                     of the ray generation off / on; the image depends on neither */
} vr_render_settings;

/* CpuRenderTarget::new + clear (render/target.rs:90-131,284-290): allocates the zeroed W*H RGBA f32
 * accumulation buffer on the device. */
int32_t vr_render_begin(vr_scene* scene, uint32_t width, uint32_t height, const vr_render_settings* settings,
                        vr_render** out);
int32_t vr_render_end(vr_render* render);
/* target.clear(): zero the accumulation buffer and the sample counter. */
int32_t vr_render_clear(vr_render* render);

/* iterative_render(target, scene, settings, samples) (iterative.rs:11-55): for every pixel draw
 * `samples` more camera samples, sum them, scale by 1/total_samples, add into the accumulation
 * buffer (alpha += 1 per call). Blocking. */
int32_t vr_render_accumulate(vr_render* render, uint32_t samples);
/* RenderAction::Cancel (render/renderer.rs:101-106): thread-safe; the running accumulate returns
 * VR_ERR_CANCELLED after the wavefront batch in flight, leaving whole samples in the buffer. Acts on the
 * accumulate that is running when it is called: with none running it does nothing (no latch carries over
 * into the next vr_render_accumulate). */
int32_t vr_render_cancel(vr_render* render);

typedef struct vr_stats {
    uint32_t samples_done;   /* RendererStats.samples.0 (renderer.rs:18-23) */
    uint32_t total_samples;  /* RendererStats.samples.1 */
    uint64_t camera_samples; /* W*H*samples_done */
    uint64_t ray_segments;   /* scene.hit calls (tracer.rs:29), counted on the device */
    double seconds;          /* host wall time spent inside vr_render_accumulate */
    double device_ms;        /* device time of all accumulate work (CUDA events on the context stream) */
    double trace_ms;         /* device time of the closest-hit kernel launches only */
    uint64_t trace_launches;
    uint64_t kernel_launches; /* every kernel this library launched for this render */
} vr_stats;
int32_t vr_render_stats(vr_render* render, vr_stats* out);

/* Copy the accumulation buffer (W*H*4 f32: partial sums already divided by total_samples,
 * target.rs:22-51) to host memory. */
int32_t vr_render_read_accum(vr_render* render, float* rgba);
/* Device pointer of the accumulation buffer (W*H*4 f32) for collectives (ncclReduce over NVLink)
 * issued by the caller on the context stream. */
int32_t vr_render_accum_device_ptr(vr_render* render, void** device_ptr);

/* ---- multi-GPU: peer-memory reduce (no counterpart in the reference, which is single-device) ----------
 * One process per GPU. Every rank exports its accumulation buffer as a CUDA IPC handle (64 bytes, exchanged by
 * the caller); the root maps the peers' buffers over NVLink and sums them with peer loads inside its own kernel,
 * in list order (deterministic, unlike a ring / tree collective). The caller must order the calls: peers finish
 * vr_render_accumulate (it is blocking) before the root reduces, and do not clear before the root is done. */
#define VR_IPC_HANDLE_BYTES 64
int32_t vr_render_export_accum(vr_render* render, uint8_t handle[VR_IPC_HANDLE_BYTES]);
/* accum(root) += sum of peers (in place, on the device). */
int32_t vr_render_reduce_peers(vr_render* render, const uint8_t* peer_handles, uint32_t n_peers);
/* Fused reduce + PostProcessingPass::render: rgba_out = tonemap(scale * (accum(root) + sum of peers)) in ONE kernel,
 * the root's accumulation buffer is left untouched. */
int32_t vr_render_resolve_peers(vr_render* render, const uint8_t* peer_handles, uint32_t n_peers, float scale,
                                float gamma, float exposure, int32_t tonemap, float* rgba_out);
/* The same two operations on raw device pointers (peers living in this process, e.g. several renders or devices
 * with peer access enabled). */
int32_t vr_render_reduce_peer_ptrs(vr_render* render, void* const* peer_accum, uint32_t n_peers);
int32_t vr_render_resolve_peer_ptrs(vr_render* render, void* const* peer_accum, uint32_t n_peers, float scale,
                                    float gamma, float exposure, int32_t tonemap, float* rgba_out);

/* PostProcessingPass::render(src, dst, PostProcessingData{scale, gamma, exposure, tonemap})
 * (render/post_process.rs:43-86, shaders/post_process.glsl, shaders/tonemapping.glsl) over the whole
 * W*H target (the reference dispatch is hard-wired to 1024x1024, post_process.rs:74).
 * tonemap: 0 None, 1 ACES, 2 Reinhard, 3 Filmic, 4 Uncharted2 (settings.rs:36-55).
 * rgba_out: W*H*4 f32 host buffer. */
int32_t vr_render_resolve(vr_render* render, float scale, float gamma, float exposure, int32_t tonemap,
                          float* rgba_out);

/* ---- correctness gates (not in the reference) ---------------------------------------------- */
/* Camera ray of global sample `sample` of every pixel and its closest hit. surface/prim: W*H u32
 * (0xFFFFFFFF = miss), prim = triangle index within the mesh; t: W*H f32 (inf on a miss). */
int32_t vr_debug_trace_primary(vr_render* render, uint32_t sample, uint32_t* surface, uint32_t* prim, float* t);
/* Closest hit of n arbitrary rays (directions are normalised like Ray::new, util/ray.rs:12-17). */
int32_t vr_debug_trace_rays(vr_scene* scene, uint64_t n, const float* origins, const float* directions,
                            uint32_t* surface, uint32_t* prim, float* t);
/* Radiance of single camera samples: out[3*i..] = trace_ray for (pixel[i], sample[i]). */
int32_t vr_debug_sample_radiance(vr_render* render, uint64_t n, const uint32_t* pixel, const uint32_t* sample,
                                 float* out);
/* Tie rank of every triangle of a mesh surface in the reference's in-order leaf sequence
 * (core/bvh.rs:48-130,171; core/mesh.rs:131). */
/* Host-only: the in-order leaf sequence of the reference's median-split tree (core/bvh.rs:48-130) over n boxes
 * (6 floats each: min xyz, max xyz) — the order the tie ranks are taken from. order[i] = item at position i. */
int32_t vr_debug_reference_leaf_order(const float* boxes6, uint64_t n, uint32_t* order);
/* Host-only: flattens one mesh the way vr_scene_commit does (reference-order tie ranks + SAH BVH2 + packed records,
 * csrc/scene_build.cpp) and returns the FNV-1a digest of the bytes the device would receive (nodes, intersection
 * records, shading records), so the shipped builder is pinned without a GPU. uvs / normals may be NULL (zeros). */
int32_t vr_debug_flatten_mesh_digest(const float* positions, const float* uvs, const float* normals, uint32_t n_vertices,
                                     const uint32_t* indices, uint32_t n_indices, uint64_t* digest, uint32_t* n_nodes,
                                     uint32_t* bvh_depth, double* flatten_ms);
int32_t vr_debug_tie_ranks(vr_scene* scene, uint32_t surface, uint32_t* out, uint32_t n);
/* Device-side evaluations of texture / environment lookups and the samplers, for unit parity. */
int32_t vr_debug_texture_sample(vr_scene* scene, uint32_t texture, uint64_t n, const float* uv, float* rgb);
int32_t vr_debug_environment_sample(vr_scene* scene, uint64_t n, const float* directions, float* rgb);
int32_t vr_debug_rng_draws(vr_context* ctx, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, uint32_t* out);
int32_t vr_debug_unit_sphere(vr_context* ctx, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif /* VOIDRAY_CUDA_H */
