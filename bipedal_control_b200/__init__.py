"""B200-native batched bipedal MPC (hot path of zitongbai/bipedal_control behind an MPC_BASE-shaped C ABI)."""
from .mpc import BatchedMpcMrtInterface, BmpcError, load_library, trot_schedule, DEFAULT_MODELS  # noqa: F401
