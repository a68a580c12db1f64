"""B200-native BiCodec detokenize (Spark-TTS hot path): semantic + global tokens -> 16 kHz waveform.

Public surface mirrors the reference (paths relative to /root/reference):
  * ``BiCodecTokenizer.detokenize(global_tokens, semantic_tokens)``  sparktts/models/audio_tokenizer.py:132-146
  * ``BiCodec.detokenize(semantic_tokens, global_tokens)``           sparktts/models/bicodec.py:171-189
  * ONNX ``bicodec_vocoder`` I/O contract                            export_sparktts_onnx.py:767-867
"""
from .config import BiCodecConfig, load_bicodec_yaml  # noqa: F401
from .bicodec import BiCodec  # noqa: F401,E402
from .audio_tokenizer import BiCodecTokenizer  # noqa: F401,E402
from .onnx_contract import VocoderSession  # noqa: F401,E402
from . import token_feed  # noqa: F401,E402
