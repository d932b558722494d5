"""Worker of tests/test_host_logic.py::test_two_rank_gloo_sharding (run under
torch.distributed.run, backend gloo, CPU only)."""
import sys

import numpy as np
import torch
import torch.distributed as dist

from ppgs_b200 import data, parallel


def main():
    out, files = sys.argv[1], sys.argv[2:]
    rank, world = parallel.init('gloo')
    assert world == 2
    loader = data.loader(files, num_workers=0, max_frames=300, shard=(rank, world))
    mine = [name for _, _, names in loader for name in names]
    with open(f'{out}/rank{rank}.txt', 'w') as f:
        f.write('\n'.join(mine))
    # the one collective of the path: rank 0's packed weight bytes to everyone
    blob = torch.arange(4096, dtype=torch.uint8) if rank == 0 else torch.zeros(4096, dtype=torch.uint8)
    dist.broadcast(blob, src=0)
    np.save(f'{out}/blob{rank}.npy', blob.numpy())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
