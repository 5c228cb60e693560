"""A third opinion on the radiance recursion: render/iterative.rs:25-42 -> core/camera.rs:26-82 -> core/tracer.rs:19-56 ->
core/mesh.rs:144-189 -> util/ray.rs:34-49 -> voidray_common/src/simple.rs:103-132 restated in numpy f32, one camera
sample at a time, with a brute-force closest hit — independent of the oracle's C++ (different language, no tree, no
shared helper) apart from reading the same reference sources. Draws come from a Python Philox4x32-10 pinned by the
Random123 known answers, laid out as the oracle documents (counter = (pixel, sample, block, 0), key = seed). On the
cube (12 triangles: Mesh::hit takes the BVH path, whose result is the smallest t) under a uniform environment no libm
function with differing implementations is on the path except atan2 in the 30-degree normal test, whose inputs are
far from the threshold on a cube, so the per-sample radiance must agree with the oracle bit for bit."""
import numpy as np

from voidray_b200.assets import asset_path, load_obj
from voidray_b200.scene import Camera, Environments, Materials, RenderSettings, Scene

F = np.float32
M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return [c0, c1, c2, c3]


class Rng:
    def __init__(self, seed, pixel, sample):
        self.key = (seed & M32, seed >> 32)
        self.pixel, self.sample, self.n, self.block, self.buf = pixel, sample, 0, None, None

    def u32(self):
        b = self.n >> 2
        if b != self.block:
            self.buf, self.block = philox4x32_10((self.pixel, self.sample, b, 0), self.key), b
        v = self.buf[self.n & 3]
        self.n += 1
        return v

    def v01(self):  # rand 0.8.5: f32::from_bits((u32 >> 9) | 0x3f80_0000) - 1.0
        return np.array([(self.u32() >> 9) | 0x3F800000], np.uint32).view(F)[0] - F(1.0)

    def gen_range(self, low, high):  # UniformFloat::sample_single
        scale = F(high - low)
        while True:
            res = F(F(self.v01() * scale) + low)
            if res < high:
                return res

    def unit_sphere(self):  # rand_distr 0.4.3 UnitSphere (Marsaglia)
        while True:
            x1 = F(F(self.v01() * F(2.0)) + F(-1.0))
            x2 = F(F(self.v01() * F(2.0)) + F(-1.0))
            s = F(F(x1 * x1) + F(x2 * x2))
            if s >= F(1.0):
                continue
            factor = F(F(2.0) * np.sqrt(F(F(1.0) - s)))
            return np.array([x1 * factor, x2 * factor, F(F(1.0) - F(F(2.0) * s))], F)


def dot(a, b):  # cgmath 0.18: mul_element_wise(a, b).sum() = (x + y) + z
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def cross(a, b):
    return np.array([F(a[1] * b[2]) - F(a[2] * b[1]), F(a[2] * b[0]) - F(a[0] * b[2]), F(a[0] * b[1]) - F(a[1] * b[0])], F)


def normalize(v):  # v * (1 / magnitude(v))
    return (v * F(F(1.0) / np.sqrt(dot(v, v)))).astype(F)


class NumpyTracer:
    def __init__(self, mesh, albedo, env, eye, center, up, fov, w, h, max_bounces, clamp, seed):
        self.pos = mesh.positions.astype(F)
        self.nrm = mesh.normals.astype(F)
        self.tri = mesh.indices.reshape(-1, 3)
        p0, p1, p2 = (self.pos[self.tri[:, k]] for k in range(3))
        self.ng = np.stack([normalize(cross(p2[i] - p1[i], p0[i] - p1[i])) for i in range(len(self.tri))])  # mesh.rs:80-84
        self.albedo, self.env = np.array(albedo, F), np.array(env, F)
        eye, center, up = np.array(eye, F), np.array(center, F), np.array(up, F)
        self.direction = normalize(center - eye)                                    # camera.rs:27
        self.up = normalize(up - (dot(up, self.direction) * self.direction).astype(F))  # camera.rs:28
        self.eye = eye
        self.d = F(F(1.0) / np.tan(F(F(fov) / F(2.0)), dtype=F))                      # camera.rs:39 (tan().recip())
        self.right = normalize(cross(self.direction, self.up))                       # camera.rs:40
        self.w, self.h, self.max_bounces, self.clamp, self.seed = w, h, max_bounces, F(clamp), seed

    def hit(self, o, d):  # Triangle::hit over every triangle; smallest t > 1e-5
        best = None
        eps = F(0.00001)
        for i, (a, b, c) in enumerate(self.tri):
            v0, v1, v2 = self.pos[a], self.pos[b], self.pos[c]
            e1, e2 = (v1 - v0).astype(F), (v2 - v0).astype(F)
            hh = cross(d, e2)
            aa = dot(e1, hh)
            if -eps < aa < eps:
                continue
            f = F(F(1.0) / aa)
            s = (o - v0).astype(F)
            u = F(f * dot(s, hh))
            if u < 0 or u > 1:
                continue
            q = cross(s, e1)
            v = F(f * dot(d, q))
            if v < 0 or F(u + v) > 1:
                continue
            t = F(f * dot(e2, q))
            if not t > eps:
                continue
            if best is None or t < best[0]:
                best = (t, i, u, v)
            elif t == best[0]:
                return "tie"
        return best

    def trace(self, o, d, depth, rng):  # tracer.rs:19-56
        if not depth < self.max_bounces:
            return np.zeros(3, F)
        h = self.hit(o, d)
        if h == "tie":
            raise ArithmeticError
        if h is None:
            return self.env.copy()
        t, i, u, v = h
        a, b, c = self.tri[i]
        w = F(F(F(1.0) - u) - v)
        n = ((u * self.nrm[b]).astype(F) + (v * self.nrm[c]).astype(F)).astype(F) + (w * self.nrm[a]).astype(F)  # mesh.rs:176
        n = n.astype(F)
        cr = cross(n, self.ng[i])
        if F(np.arctan2(np.sqrt(dot(cr, cr)), dot(n, self.ng[i]), dtype=F)) > F(F(F(30.0) * F(np.pi)) / F(180.0)):
            n = self.ng[i]
        point = (o + (d * t).astype(F)).astype(F)          # Ray::at
        normal = n if dot(d, n) < 0 else (-n).astype(F)     # HitRecord::new
        sd = (normal + rng.unit_sphere()).astype(F)         # simple.rs:116
        if abs(sd[0]) < F(1e-8) and abs(sd[1]) < F(1e-8) and abs(sd[2]) < F(1e-8):
            sd = normal
        inner = self.trace(point, normalize(sd), depth + 1, rng)
        delta = (self.albedo * inner).astype(F)
        return (np.zeros(3, F) + np.minimum(delta, self.clamp)).astype(F)  # color += delta.clamp(max): channel-wise min

    def sample(self, pixel, sample):  # iterative.rs:25-42 with y = index / width, then camera.rs:69-82
        rng = Rng(self.seed, pixel, sample)
        W, H = self.w, self.h
        px, py = pixel % W, pixel // W
        dd = F(max(W, H))
        x = F(F(F(2 * px + 1) - F(W)) / dd)
        y = F(F(F(2 * (H - py) - 1) - F(H)) / dd)
        dx = rng.gen_range(F(F(-1.0) / dd), F(F(1.0) / dd))
        dy = rng.gen_range(F(F(-1.0) / dd), F(F(1.0) / dd))
        cx, cy = F(x + dx), F(y + dy)
        new_dir = ((self.d * self.direction).astype(F) + (cx * self.right).astype(F)).astype(F) + (cy * self.up).astype(F)
        return self.trace(self.eye, normalize(normalize(new_dir.astype(F))), 0, rng)  # camera.rs:81, then Ray::new


def test_python_philox_known_answers():
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox4x32_10((M32,) * 4, (M32, M32)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_numpy_tracer_matches_the_oracle_sample_for_sample(oracle):
    w, h, bounces, clamp, seed = 64, 48, 6, 3.0, 0x5EED0001
    eye, center, up, fov = (2.5, 1.8, -4.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 0.6
    albedo, env = (0.6, 0.5, 0.4), (0.7, 0.8, 0.9)
    mesh = load_obj(asset_path("cube.obj"))
    scene = Scene.empty()
    scene.add_object(scene.add_material(Materials.lambertian(albedo)), scene.add_mesh(mesh))
    scene.environment = Environments.uniform(env)
    scene.camera = Camera.look_at(eye, center, up, fov)
    rs = RenderSettings(total_samples=16, max_bounces=bounces, firefly_clamp=clamp, seed=seed)
    rng = np.random.default_rng(1)
    px = rng.integers(0, w * h, 400).astype(np.uint32)
    sm = rng.integers(0, 16, 400).astype(np.uint32)
    ref = oracle.OracleScene(scene).sample_radiance(w, h, rs, px, sm)
    tracer = NumpyTracer(mesh, albedo, env, eye, center, up, fov, w, h, bounces, clamp, seed)
    compared = hits = 0
    with np.errstate(all="ignore"):
        for k in range(len(px)):
            try:
                got = tracer.sample(int(px[k]), int(sm[k]))
            except ArithmeticError:
                continue  # an exact tie between two triangles (a ray through a cube edge): the tie rule is tested elsewhere
            compared += 1
            hits += int(np.any(got != np.array(env, F)))
            assert np.array_equal(got.view(np.uint32), ref[k].view(np.uint32)), (k, got, ref[k])
    assert compared > 380 and hits > 60
