import os, torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((world, 1024), dtype=torch.float64, device='cuda')
    t.fill_(-1.0)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    ptrs = list(hdl.buffer_ptrs)
    print(rank, 'rendezvous ok', [hex(p) for p in ptrs], 'rank', hdl.rank, 'world', hdl.world_size, flush=True)
    # write my rank id into row `rank` of every peer's buffer via peer views
    for p in range(world):
        peer = hdl.get_buffer(p, (world, 1024), torch.float64)
        peer[rank].fill_(float(rank))
    torch.cuda.synchronize()
    hdl.barrier()
    torch.cuda.synchronize()
    print(rank, 'rows', t[:, 0].tolist(), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, 'SYMM FAILED', repr(e), flush=True)
dist.barrier(); dist.destroy_process_group()
