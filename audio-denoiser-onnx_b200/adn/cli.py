"""`python -m adn.cli MODEL.adn IN.wav OUT.wav`: the demo flow of the reference's `Inference_*_ONNX.py` scripts
(load wav -> fixed windows -> run -> concatenate -> write wav -> print RTF, `GTCRN/Inference_GTCRN_ONNX.py:268-344`)
on the B200 path, with all windows of the file in ONE batched run.  Models with several outputs
(MossFormer2-SS) write OUT_0.wav, OUT_1.wav.  The wav must already be at the model's input sample rate (the
reference resamples with ffmpeg through pydub; no resampler is applied here)."""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

from . import chunker, wavio
from .ort_shim import InferenceSession


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) != 3:
        print(__doc__)
        return 2
    model, src, dst = argv
    sess = InferenceSession(model)
    md = sess.get_modelmeta().custom_metadata_map
    i0 = sess.get_inputs()[0]
    audio, sr = wavio.read_wav(src)
    want_sr = int(md.get("in_sample_rate", sr))
    if sr != want_sr:
        raise SystemExit(f"{src}: sample rate {sr} != model input rate {want_sr}")
    if "int16" not in i0.type:
        raise SystemExit("the CLI drives INT16-in / INT16-out model files (the reference's default I/O dtype)")
    chans = i0.shape[-2]
    x = wavio.to_mono(audio) if chans == 1 else chunker.match_channels(audio, chans)
    n_out = len(sess.get_outputs())
    t0 = time.time()
    if n_out > 1:
        ys = chunker.separate(sess, x)
    else:
        ys = [chunker.denoise(sess, x)]
    dt = time.time() - t0
    out_sr = int(md.get("out_sample_rate", sr))
    dst = Path(dst)
    for k, y in enumerate(ys):
        path = dst if n_out == 1 else dst.with_name(f"{dst.stem}_{k}{dst.suffix}")
        wavio.write_wav(path, np.asarray(y, dtype=np.int16), out_sr)
        print(f"wrote {path}")
    dur = x.shape[-1] / float(sr)
    print(f"RTF: {dt / dur:.6f}  ({dur:.2f} s of audio in {dt * 1e3:.1f} ms)")        # :341-344
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
