#!/usr/bin/env python
"""SURVEY §8f row 2 as a workload: a DIRECTORY of small DDS files (256x256 BC1 / BC2 with full mip chains, legacy and DX10
headers) read into one page-locked pool and pushed through dltdds_transform_bundle_batch / dltdds_untransform_batch —
what the reference's CLI does with one rayon task per file — next to the same files through the single-file calls.
Manual (default settings) and auto (LTU estimator, search per file) bundles.  Prints one JSON line per case."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import dds_fixtures as fx  # noqa: E402
import dxt_lossless_transform_b200 as dlt  # noqa: E402
from dxt_lossless_transform_b200 import file_formats as ff  # noqa: E402
from dxt_lossless_transform_b200 import synth  # noqa: E402


def main():
    torch.cuda.set_device(0)
    nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    def dx10_file(dxgi: int, legacy: np.ndarray) -> np.ndarray:
        """The same texture behind a DX10 header (148 bytes: the payload is then only 4-byte aligned in the file)."""
        import struct
        d = fx.header_base(256, 256, 9, True)
        d[0x54:0x58] = b"DX10"
        struct.pack_into("<I", d, 0x50, 0x4)
        struct.pack_into("<I", d, 0x80, dxgi)      # DXGI_FORMAT_BC1_UNORM = 71, BC2_UNORM = 74
        struct.pack_into("<II", d, 0x84, 3, 0)      # resourceDimension = TEXTURE2D
        struct.pack_into("<I", d, 0x90, 1)          # arraySize
        return np.frombuffer(bytes(d) + legacy[128:].tobytes(), np.uint8).copy()

    templates = []
    for fmt, bcn, dxgi in ((ff.DdsFormat.BC1, 1, 71), (ff.DdsFormat.BC2, 2, 74)):
        legacy = fx.make_dds(int(fmt), 256, 256, mips=9)
        templates.append((bcn, legacy))
        dx = dx10_file(dxgi, legacy)
        info = ff.parse_dds(dx)
        assert info is not None and info.data_offset == 148 and int(info.format) == int(fmt), info
        templates.append((bcn, dx))
    handler = ff.DdsHandler()
    from dxt_lossless_transform_b200 import _native as N
    all_templates = templates
    for kind, templates in (("legacy FourCC headers (payload at 128)", all_templates[0::2]), ("DX10 headers (payload at 148)", all_templates[1::2]),
                            ("mixed", all_templates)):
        sizes = [len(t[1]) for t in templates]
        stride = (max(sizes) + 255) // 256 * 256
        pin_in, pin_out, pin_back = (dlt.alloc_pinned(stride * nfiles) for _ in range(3))
        pairs, back_pairs, total = [], [], 0
        for i in range(nfiles):
            bcn, tmpl = templates[i % len(templates)]
            n = len(tmpl)
            info = ff.parse_dds(tmpl)
            buf = pin_in.array[i * stride:i * stride + n]
            buf[:] = tmpl
            nb = info.data_length // (8 if bcn == 1 else 16)
            buf[info.data_offset:info.data_offset + info.data_length] = synth.texture_blocks(bcn, nb, seed=i % 97)
            pairs.append((buf, pin_out.array[i * stride:i * stride + n]))
            back_pairs.append((pin_out.array[i * stride:i * stride + n], pin_back.array[i * stride:i * stride + n]))
            total += n
        # the C entry points are timed directly (marshalling thousands of buffers through ctypes is Python's cost)
        arr, _keep = handler._files(pairs)
        back_arr, _keep2 = handler._files(back_pairs)
        res = (N.DltffResult * nfiles)()

        def timed(fn, reps=3):
            fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            return (time.perf_counter() - t0) / reps

        manual = ff.TransformBundle.default_all()
        ltu = dlt.LosslessTransformUtilsSizeEstimation()
        auto = ff.TransformBundle.new().with_bc1_auto(dlt.Bc1AutoTransformBuilder(ltu)).with_bc2_auto(dlt.Bc2AutoTransformBuilder(ltu))
        for label, bundle in (("manual (default settings)", manual), ("auto (LTU estimator, search per file)", auto)):
            def batch():
                assert N.lib().dltdds_transform_bundle_batch(arr, nfiles, bundle._h, res, None, 0) == 0
            dt = timed(batch)
            assert all(ff.error_from_result(r) is None for r in res)
            dtu = timed(lambda: N.lib().dltdds_untransform_batch(back_arr, nfiles, res, None, 0))
            ok = all(np.array_equal(b[1], p[0]) for b, p in zip(back_pairs[:50], pairs[:50]))
            sub = pairs[:200]
            dts = timed(lambda: [handler.transform_bundle(i, o, bundle) for i, o in sub], reps=1) * len(pairs) / len(sub)
            print(json.dumps({"files": kind, "bundle": label, "count": nfiles, "bytes": total, "batch_ms": dt * 1e3,
                              "files_per_s": nfiles / dt, "batch_gbs": total / dt / 1e9, "untransform_batch_ms": dtu * 1e3,
                              "round_trip_ok_first50": bool(ok), "single_file_calls_ms_extrapolated": dts * 1e3}), flush=True)
        for p in (pin_in, pin_out, pin_back):
            p.free()


if __name__ == "__main__":
    main()
