from ..inference.functions import (append_default_compiler_flags, compile, compile_layout_transformer, compile_wlo,  # noqa: F401,A004
                                   shard_checkpoint, trace)
