"""vrfs-b200: B200-native batched VRF engine for the hot path of ark-ec-vrfs."""
from ._lib import VrfsError, LIB_PATH  # noqa: F401
from .engine import (Engine, MultiEngine, host_buffer, host_copy, BANDERSNATCH, ED25519, P256, BANDERSNATCH_SW, JUBJUB, BABYJUBJUB, pack_var,  # noqa: F401
                     ITEM_OK, ITEM_VERIFICATION_FAILURE, ITEM_INVALID_DATA)
