"""Small driver for ncu: a few K=64 rollout launches of config 4 (computer vs computer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pikazoo_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 8
env = pikazoo_b200.PikaVecEnv(n, seed=4040, winning_score=15, serve="winner", is_player1_computer=True,
                              is_player2_computer=True)
env.reset()
for _ in range(warm + 2):
    env.rollout(64)
torch.cuda.synchronize()
print("ok", env.stats_dict())
