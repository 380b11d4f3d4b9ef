"""Where does the end-to-end time of hx.price_distributed go?  (torchrun, N ranks)"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import hestonexotics_b200 as hx
from hestonexotics_b200 import _lib
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
os.environ["NCCL_DEBUG"] = "WARN"; os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = _lib.load(); _lib.check(lib.hexo_gpu_init(lr))
A = hx.HQEAnderson(hx.AAsianCallNonAdaptive); p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
ch = [hx.OptionsChain.from_strikes(1.0, [100.0])]
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
ns = int(lib.hexo_gpu_default_streams(n, 1, world))
def call():
    return hx.price_distributed(A, p, 100.0, ch, n, 1, 1024, seed=1, n_streams=ns)
call(); dist.barrier(); torch.cuda.synchronize()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    r = call()
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"rank {rank} it {it}: wall {1e3*(t1-t0):8.2f} ms  device {e0.elapsed_time(e1):8.2f} ms", flush=True)
# the same shard through a prepared plan
rq = hx.pricing._Request(A, p, 100.0, ch, n, 1, 1024, 1, "f32", ns)
b, c = hx.shard_range(ns, rank, world)
plan = C.c_void_p(); _lib.check(lib.hexo_gpu_plan_create(C.byref(rq.req), b, c, C.byref(plan)))
sums = torch.zeros(2, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()
for it in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter(); e0.record()
    _lib.check(lib.hexo_gpu_plan_launch(plan, C.c_void_p(sums.data_ptr()), C.c_void_p(st.cuda_stream)))
    e1.record(); dist.all_reduce(sums); e2.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"rank {rank} plan it {it}: wall {1e3*(t1-t0):8.2f} ms  kernel {e0.elapsed_time(e1):8.2f}  +allreduce {e1.elapsed_time(e2):8.2f}", flush=True)
dist.destroy_process_group()
