import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
if rank == 0:
    print("symm_mem attrs:", [a for a in dir(symm_mem) if not a.startswith("__")][:60])
try:
    t = symm_mem.empty(1024, dtype=torch.float32, device=torch.device("cuda", local))
    h = symm_mem.rendezvous(t, dist.group.WORLD)
    if rank == 0:
        print("handle attrs:", [a for a in dir(h) if not a.startswith("_")])
        print("multicast_ptr", getattr(h, "multicast_ptr", None)); print("buffer_ptrs", h.buffer_ptrs, "signal_pad_ptrs", h.signal_pad_ptrs, "rank", h.rank, "world", h.world_size)
    t.fill_(rank + 1)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (1024,), torch.float32)
    print(rank, "peer value", float(peer[0]))
    h.barrier()
except Exception as e:
    import traceback; traceback.print_exc()
dist.destroy_process_group()
