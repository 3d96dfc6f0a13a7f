"""mocca_envs_b200 -- B200-native batched replacement for the mocca_envs hot path.

Public surface mirrors the reference's: ``make(env_id)`` (gym.make analogue, reference mocca_envs/__init__.py),
gym-style ``reset/step/seed`` plus a batched VecEnv returning torch CUDA tensors.
"""
ENV_IDS = ["Walker3DCustomEnv-v0", "Walker3DStepperEnv-v0", "Monkey3DCustomEnv-v0", "CassieEnv-v0",
           "Child3DCustomEnv-v0", "MikeStepperEnv-v0", "Walker2DCustomEnv-v0", "Crab2DCustomEnv-v0"]


def __getattr__(name):
    if name in ("make", "Walker3DCustomVecEnv", "Walker3DCustomEnv", "Walker3DStepperVecEnv", "Walker3DStepperEnv",
                "Monkey3DCustomVecEnv", "Monkey3DCustomEnv", "CassieVecEnv", "CassieEnv", "Child3DCustomVecEnv",
                "Child3DCustomEnv", "MikeStepperVecEnv", "MikeStepperEnv", "Walker2DCustomVecEnv", "Walker2DCustomEnv",
                "Crab2DCustomVecEnv", "Crab2DCustomEnv"):
        from . import vec_env

        return getattr(vec_env, name)
    raise AttributeError(name)
