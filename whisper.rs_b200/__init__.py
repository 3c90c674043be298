"""whisper.rs_b200 -- B200-native (sm_100a) Whisper hot path behind the entry points of
szuwgh/whisper.rs (`WhisperContext::new`, `whisper_pcm_to_mel`, `whisper_encode`, and the decode
call the reference leaves unimplemented).

The directory name carries a dot, so it is loaded through `__graft_entry__.load_package()`
(importlib, registered as `whisper_rs_b200`) rather than a plain `import`.

  ggml_file  -- ggml-v1 model files: writer (random-init fixtures) + host reader
  synth      -- seeded synthetic 16 kHz PCM
  shard      -- segment -> GPU assignment and the final gather (the path's only collective)
  cabi       -- ctypes binding of csrc/libwhisper_b200.so (the C-ABI drop-in boundary)
  api        -- host-side mirror of the reference's functions over the C-ABI
"""
from . import ggml_file, shard, synth  # noqa: F401
