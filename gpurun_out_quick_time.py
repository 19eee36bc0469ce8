import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/sparse-voxel-octrees_b200')
import numpy as np, torch, pysvo
tree = pysvo.VoxelOctree('/root/repo/tests/golden/XYZRGB-Dragon.oct')
for (W,H) in [(1280,720),(3840,2160)]:
    buf = torch.zeros(H*W, dtype=torch.int32, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for cam_args in [(0,0,1.0),(20,135,0.5)]:
        cam = pysvo.orbit_camera(*cam_args)
        for fl in (0,1):
            st = tree.render_frame_device(cam, W, H, buf.data_ptr(), flavour=fl, stream=s, want_stats=True)
            for _ in range(3): tree.render_frame_device(cam, W, H, buf.data_ptr(), flavour=fl, stream=s)
            torch.cuda.synchronize()
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): tree.render_frame_device(cam, W, H, buf.data_ptr(), flavour=fl, stream=s)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)/20
            print(W,H,cam_args,'flavour',fl,'ms/frame %.4f'%ms, 'rays', st.rays, 'Mrays/s %.1f'%(st.rays/ms/1e3))
