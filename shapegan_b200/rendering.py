"""Device-resident sphere tracing on the fused single-latent SDFNet kernel -- the inner loops of rendering/raymarching.py
(`render_image` :106-120, `get_shadows` :48-61) without their per-iteration host round trips.

The reference advances `points[indices]` by the clamped SDF once per iteration: gather -> `evaluate_in_batches` (chunks of 100 000 with
the latent `.repeat`ed) -> clamp -> scatter -> two boolean-mask compactions, ~10 kernel launches and a host sync per iteration, up to
1000 iterations per image.  Here ONE launch per iteration does all of it (sg_sdfnet_infer in trace mode): the kernel reads the rays of
the current list, evaluates the MLP with the latent folded into the bias, advances the points in place, marks hits, drops misses and
appends the survivors to the next list with its own device-side count; the host only looks at the count every `check_every` steps."""
import torch

from . import raw, sdf_ops


class SphereTracer:
    """rays: points [N,3] fp32 (advanced in place), directions [N,3] fp32, both on the SDFNet's device.

    hit  = 0 < clamp(sdf + sdf_offset) < threshold                     (raymarching.py:112, :55)
    miss = |p| > radius after the step (miss_y: p.y > radius, the shadow rays of :59)
    Rays still alive when the iterations run out count as hits (:121, :64)."""

    def __init__(self, sdf_net, latent_code, points, directions, indices=None, clamp=0.02, threshold=0.0005, radius=1.0, sdf_offset=0.0,
                 miss_y=False):
        if not sdf_ops.folded_enabled(sdf_net.latent_code_size):
            raise RuntimeError('SphereTracer needs the fused bf16 SDFNet kernel (precision bf16, latent_code_size 128)')
        self.net = sdf_net
        dev = points.device
        self.points = points.contiguous()
        assert self.points.data_ptr() == points.data_ptr(), 'points must be contiguous: they are advanced in place'
        self.dirs = directions.contiguous().float()
        n = points.shape[0]
        self.n = n
        self.hit = torch.zeros(n, dtype=torch.uint8, device=dev)
        first = torch.arange(n, dtype=torch.int32, device=dev) if indices is None else indices.to(device=dev, dtype=torch.int32).contiguous()
        self.lists = [first.clone() if first.shape[0] == n else torch.cat((first, torch.zeros(n - first.shape[0], dtype=torch.int32, device=dev))),
                      torch.empty(n, dtype=torch.int32, device=dev)]
        self.cur = 0
        self.count0 = int(first.shape[0])
        self.params = dict(sdf_offset=float(sdf_offset), clamp=float(clamp), threshold=float(threshold), radius=float(radius), miss_y=bool(miss_y))
        self.img, self.aux = sdf_ops.folded_weights(sdf_net._params(), latent_code.to(dev))
        self.steps_done = 0

    def run(self, iterations, check_every=25):
        """Trace up to `iterations` steps; returns the uint8 hit mask [N] (device).  Stops early when fewer than 2 rays are alive
        (the reference's `if indices.shape[0] < 2: break`, evaluated every `check_every` steps)."""
        dev = self.points.device
        if self.count0 == 0:
            return self.hit
        # counts[i] = rays alive before step i: one zero-filled array, step i reads counts[i] and appends into counts[i + 1]
        counts = torch.zeros(iterations + 1, dtype=torch.int32, device=dev)
        counts[0] = self.count0
        alive, done = self.count0, 0
        with torch.no_grad():
            for i in range(iterations):
                nxt = 1 - self.cur
                trace = dict(points=self.points, dirs=self.dirs, hit=self.hit, next_index=self.lists[nxt], next_count=counts[i + 1:], **self.params)
                raw.sdfnet_infer(self.img, self.aux, n=alive, ray_index=self.lists[self.cur], n_ptr=counts[i:], trace=trace)
                self.cur, done = nxt, i + 1
                if done % check_every == 0 or done == iterations:
                    alive = int(counts[done].item())            # also tightens the grid bound of the following launches
                    if alive < 2:
                        break
            # rays that never terminated count as hits (raymarching.py:121 / :64)
            if alive > 0:
                self.hit[self.lists[self.cur][:alive].long()] = 1
        self.steps_done += done
        self.count0 = 0
        return self.hit


def trace_camera_rays(sdf_net, latent_code, points, directions, indices, iterations=1000, threshold=0.0005, sdf_offset=0.0, radius=1.0):
    """render_image's marching loop (raymarching.py:106-121): returns the model mask; `points` holds the hit positions afterwards."""
    return SphereTracer(sdf_net, latent_code, points, directions, indices, clamp=0.02, threshold=threshold, radius=radius,
                        sdf_offset=sdf_offset).run(iterations)


def trace_shadow_rays(sdf_net, latent_code, points, directions, iterations=200, threshold=0.001, sdf_offset=0.0, radius=1.0):
    """get_shadows' loop (raymarching.py:43-64) for surface points already offset along the light direction: 1 = in shadow."""
    return SphereTracer(sdf_net, latent_code, points, directions, None, clamp=0.1, threshold=threshold, radius=radius, sdf_offset=sdf_offset,
                        miss_y=True).run(iterations)
