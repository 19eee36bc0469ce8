"""How long does the HOST take to issue one frame (ctypes + events)? Tiny frames, so the GPU never is the limit."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/sparse-voxel-octrees_b200')
import torch, pysvo
tree = pysvo.VoxelOctree('/root/repo/tests/golden/XYZRGB-Dragon.oct')
W, H = 160, 88
buf = torch.zeros(H * W, dtype=torch.int32, device='cuda')
lanes = [torch.cuda.Stream() for _ in range(4)]
cams = [pysvo.orbit_camera(0, 3.6 * k, 1.0) for k in range(100)]
flag = torch.zeros(1, device='cuda', dtype=torch.int32)
for mode in ("frame only", "frame + 2 events + wait"):
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        n = 3000
        for k in range(n):
            s = lanes[k % 4]
            with torch.cuda.stream(s):
                tree.render_frame_device(cams[k % 100], W, H, buf.data_ptr(), strips=16, flavour=1, stream=s.cuda_stream)
                if mode != "frame only":
                    e = torch.cuda.Event(); e.record(s); lanes[(k + 1) % 4].wait_event(e)
                    g = torch.cuda.Event(); g.record(s)
        t_issue = time.perf_counter() - t
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t
    print(mode, 'issue %.1f us/frame, complete %.1f us/frame' % (t_issue / n * 1e6, t_all / n * 1e6))
