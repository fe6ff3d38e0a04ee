"""Probe (2+ GPUs): does this box give torch symmetric memory a multicast (NVLS) mapping? Timings of torch's own
multimem / one-shot / two-shot all-reduce ops and NCCL on a [256, 4096] bf16 message, for orientation only."""
import os, sys, json, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world}
try:
    from cuda.bindings import driver as cu
    cu.cuInit(0)
    err, v = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, local)
    out["multicast_attr"] = int(v)
    err, v = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, local)
    out["fabric_attr"] = int(v)
except Exception as e:
    out["attr_err"] = repr(e)
gname = dist.group.WORLD.group_name
t = symm_mem.empty(256 * 4096, dtype=torch.bfloat16, device=dev)
hdl = symm_mem.rendezvous(t, gname)
out["multicast_ptr"] = int(hdl.multicast_ptr)
try:
    out["has_multicast_support"] = bool(type(hdl).has_multicast_support(dev.type, dev.index))
except Exception as e:
    out["has_multicast_support"] = repr(e)[:200]
out["offset"] = int(getattr(hdl, "offset", -1))
out["buffer_size"] = int(hdl.buffer_size); out["signal_pad_size"] = int(hdl.signal_pad_size)
out["buffer_ptrs"] = [hex(p) for p in hdl.buffer_ptrs]
out["signal_pad_ptrs"] = [hex(p) for p in hdl.signal_pad_ptrs]

def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

x = torch.randint(1, 8, (256 * 4096,), device=dev).to(torch.bfloat16)
t.copy_(x)
ref = x.clone(); dist.all_reduce(ref)
res = {}
for name in ("multimem_all_reduce_", "one_shot_all_reduce", "two_shot_all_reduce_"):
    try:
        op = getattr(torch.ops.symm_mem, name)
        t.copy_(x); torch.cuda.synchronize(); dist.barrier()
        y = op(t, "sum", gname)
        torch.cuda.synchronize()
        ok = bool(torch.equal(y.float(), ref.float()))
        g = torch.cuda.CUDAGraph()
        res[name] = {"ok": ok, "us": timeit(lambda: op(t, "sum", gname))}
    except Exception as e:
        res[name] = {"err": repr(e)[:300]}
y = x.clone()
res["nccl"] = {"us": timeit(lambda: dist.all_reduce(y))}
out["ops"] = res
if rank == 0:
    print(json.dumps(out))
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
