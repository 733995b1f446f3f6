"""Profiling driver for the ReSTIR frame (for ncu): C2 scene, 1080p, bounces 3, N frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtdx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sc = rtdx.scenes.mesh_room(n=296)
ctx = rtdx.Context(1920, 1080, bounces=3, flags=rtdx.FLAG_RESTIR)
ctx.upload_scene(sc)
for f in range(n):
    ctx.render_frame(f)
ctx.synchronize()
print(ctx.counters())
