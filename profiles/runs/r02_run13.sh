set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python profiles/r02_gbuffer_ab.py > gpurun_out/r02_gbuffer_ab.txt 2>&1; cat gpurun_out/r02_gbuffer_ab.txt
timeout 900 python -m pytest tests/test_gpu_matstage.py tests/test_gpu_texture.py tests/test_gpu_post.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests13.log 2>&1; tail -5 gpurun_out/r02_gpu_tests13.log | cut -c1-300
timeout 600 python - > gpurun_out/r02_step_cprofile.txt 2>&1 <<'PY'
import cProfile, pstats, os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench
from texir_code_b200 import models, synth, Scene
dev = torch.device('cuda:0')
sc = synth.make_room_scene(1 << 20, seed=666)
tex = torch.from_numpy(synth.make_radiance_texture(4096, seed=666))
scene = Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=dev)
mvp, cam = bench.cube_mvps(dev), torch.zeros(3, device=dev)
g = torch.randn(6, 128, 128, 3, device=dev)
mm = models.MaterialModel(None, None, sc.triangle_uvs, tex, sample_l=[32, 16], scene=scene, materials_a_res=2048,
                          materials_r_res=4096, cube_res=128, irrt=torch.rand(1024, 1024, 3) * 3.0, max_mip_level=13)
mm._warned_normals = True
def step():
    mm.zero_grad(set_to_none=True)
    mm(mvp, 0, cam, 2)["rgb"].backward(g)
for _ in range(20): step()
torch.cuda.synchronize()
# host time alone: how long the Python side needs to issue one step when the GPU is not the limit
pr = cProfile.Profile()
pr.enable()
for _ in range(200): step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
PY
head -60 gpurun_out/r02_step_cprofile.txt | cut -c1-200
