import torch, time, os, sys
sys.path.insert(0, os.getcwd())
from vfnerf_b200 import samplers
from vfnerf_b200.samplers import cpu_generator_rand_
h = torch.empty(65536*64).pin_memory()
cpu_generator_rand_(h)
print("fast rng:", samplers._FAST_RNG, "threads", torch.get_num_threads(), "cpus", os.cpu_count())
for _ in range(3):
    t=time.perf_counter(); cpu_generator_rand_(h); dt=time.perf_counter()-t
    print(f"{dt*1e3:.2f} ms for {h.numel()/1e6:.1f} M draws = {h.numel()/dt/1e6:.0f} M/s")
t=time.perf_counter(); x=torch.rand(65536*64); dt=time.perf_counter()-t; print("torch.rand", dt*1e3)
