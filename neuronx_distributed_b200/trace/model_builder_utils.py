from ..inference.model_builder_utils import (CompilationArtifacts, LayoutTransformerArtifacts, ModelBuilderConstants,  # noqa: F401
                                             ModelParamInfo, ProvidedArgInfo, TraceArtifacts, WLOArtifacts, generate_key)
