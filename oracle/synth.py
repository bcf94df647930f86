"""TEST INFRASTRUCTURE — pure-Python restatement of pk_synth_action (oracle/pika_oracle.c),
the product-defined synthetic action stream (DESIGN.md "synthetic actions"). Integer-only,
so fixtures do not depend on any library's random generator."""

M64 = (1 << 64) - 1


def synth_action(action_seed: int, global_env: int, frame: int, agent: int, n_actions: int = 18) -> int:
    z = (action_seed + 0x9E3779B97F4A7C15 * (2 * global_env + agent + 1)) & M64
    z ^= (frame * 0xD1B54A32D192ED03) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return ((z >> 32) * n_actions) >> 32


def synth_actions_numpy(action_seed: int, first_env: int, n_envs: int, frame: int, n_actions: int = 18):
    """Vectorised form: int32 array [n_envs, 2] for one call index."""
    import numpy as np

    with np.errstate(over="ignore"):
        env = np.arange(first_env, first_env + n_envs, dtype=np.uint64)
        out = np.empty((n_envs, 2), dtype=np.int32)
        for agent in (0, 1):
            z = np.uint64(action_seed) + np.uint64(0x9E3779B97F4A7C15) * (np.uint64(2) * env + np.uint64(agent + 1))
            z = z ^ (np.uint64(frame) * np.uint64(0xD1B54A32D192ED03))
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[:, agent] = (((z >> np.uint64(32)) * np.uint64(n_actions)) >> np.uint64(32)).astype(np.int32)
    return out
