for u in 1 4 8; do for g in 2 4 8; do
DM2_PEER_MC_UNROLL=$u DM2_PEER_GRID_MULT=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29521 tools/check_peer_allreduce.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('u=$u g=$g', 'p2p', d['cfg2_p2p'].get('peer_ms'), 'mc', d['cfg2_mc'].get('peer_ms') if isinstance(d['cfg2_mc'],dict) else d['cfg2_mc'], 'nccl', d['cfg2_p2p'].get('nccl_ms'), 'err', d['cfg2_p2p']['max_rel_err_vs_nccl'], d['cfg2_mc']['max_rel_err_vs_nccl'] if isinstance(d['cfg2_mc'],dict) else None)"
done; done
