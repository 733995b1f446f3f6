import os, sys, time
sys.path.insert(0, "/root/repo")
import rtdx, torch
sc = rtdx.scenes.instanced_blobs()
ctx = rtdx.Context(3840, 2160, bounces=3)
up = ctx.upload_scene(sc)
acc = {}
def t(name, fn):
    t0 = time.perf_counter(); r = fn(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
outs = [torch.empty((2160, 3840, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
for k in range(13):
    if k == 3: acc.clear(); t0 = time.perf_counter()
    t("set_instances", lambda: ctx.set_instances(up["descs"], up["props"]))
    t("set_camera", lambda: ctx.set_camera(up["camera"]))
    t("render_pass(launch)", lambda: ctx.render_pass(k, 1))
    t("wait_output(k-1)", lambda: ctx.wait_output())
    t("read_output_async", lambda: ctx.read_output_async(outs[k & 1]))
ctx.wait_output()
print("C3 one frame in flight: wall %.3f ms per frame" % ((time.perf_counter() - t0) / 10 * 1e3), {k: round(v / 10 * 1e3, 3) for k, v in acc.items()})
for name, do_set, do_read, do_cam in (("render only", False, False, False), ("set_instances + render", True, False, False), ("render + read-back", False, True, False),
                                     ("set_instances + render + read-back", True, True, False), ("set_camera + render", False, False, True), ("all", True, True, True)):
    ctx.synchronize(); ctx.wait_output()
    t0 = time.perf_counter()
    for k in range(10):
        if do_set: ctx.set_instances(up["descs"], up["props"])
        if do_cam: ctx.set_camera(up["camera"])
        ctx.render_pass(20 + k, 1)
        if do_read:
            ctx.wait_output(); ctx.read_output_async(outs[k & 1])
    ctx.wait_output(); ctx.synchronize()
    print("%-28s %.3f ms per frame" % (name, (time.perf_counter() - t0) / 10 * 1e3))
