import os, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr), "t.data_ptr", hex(t.data_ptr()), flush=True)
    print(rank, [a for a in dir(hdl) if not a.startswith("_")], flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "symm_mem failed:", type(e).__name__, e, flush=True)
dist.barrier(); dist.destroy_process_group()
