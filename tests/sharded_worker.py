"""torchrun worker of test_sharded_clip_over_two_gpus_is_bit_identical_to_one_gpu: renders this rank's block, all-gathers, rank 0 saves
what every rank holds."""
import datetime
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import sma_b200 as S
    import sma_oracle as O
    from conftest import CFG, GOLD
    import json
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local), timeout=datetime.timedelta(seconds=180))
    inv = json.load(open(os.path.join(GOLD, 'state_keys.json')))
    g = S.build_network(CFG['network_g']); me = S.build_network(CFG['network_motion_estimator'])
    g.load_state_dict(O.synthetic_state_dict(inv['net_g'], 0)); me.load_state_dict(O.synthetic_state_dict(inv['motion_estimator'], 1))
    g, me = g.eval().cuda(), me.eval().cuda()
    src, drv = O.synthetic_frames(8, seed=41)
    clip = S.make_animation_sharded(src, drv, g, me, batch=4)
    parts = [torch.empty_like(clip) for _ in range(world)]
    dist.all_gather(parts, clip)
    if rank == 0:
        torch.save({'world': world, 'rank0': parts[0].cpu(), 'rank1': parts[1].cpu()}, sys.argv[1])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
