"""Host BVH8 builder on this machine's cores (no GPU needed): phase times for the C2 soup and the C3 stand-in, default
thread placement and OMP_PROC_BIND=true.  Each configuration runs in a fresh process, because the first build of a
process is the one a renderer pays (thread creation, first-touch page faults).
    python tools/bench_host_builder.py [threads ...]"""
import os
import subprocess
import sys

CHILD = r"""
import sys, time
sys.path.insert(0, %r)
from gatling_b200 import gi, scenes
which = sys.argv[1]
d = scenes.random_triangles(1000000, 64, 64, 1) if which == "c2" else scenes.kitchen_standin(width=64, height=64, spp=1)
d.dome_light = None
for rep in range(2):
    t = time.time(); b = gi.BuiltScene(d); n, tr = gi.get_bvh(b.scene); dt = time.time() - t; b.close()
    print("%%s build %%d: %%.0f ms wall, %%d nodes, %%d triangles" %% (which, rep, dt * 1e3, len(n) // 80, len(tr) // 12), flush=True)
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

threads = [int(a) for a in sys.argv[1:]] or [os.cpu_count() or 8]
for which in ("c2", "c3"):
    for th in threads:
        for bind in (None, "true"):
            env = dict(os.environ, OMP_NUM_THREADS=str(th), GTLC_BUILD_TIMING="1")
            if bind:
                env["OMP_PROC_BIND"] = bind
            print("== %s, %d threads, OMP_PROC_BIND=%s" % (which, th, bind), flush=True)
            subprocess.run([sys.executable, "-c", CHILD, which], env=env)
