"""SURVEY.md 8(f1) feasibility probe: is NVDEC usable from this container?  (run on the GPU box; writes a text report)

Checks, in order: is libnvcuvid (the driver's video decode library) present and loadable; does cuvidGetDecoderCaps
report H.264 4:2:0 8-bit decode on this GPU; can the demux side be had without ffmpeg headers (OpenCV's FFmpeg backend
with CAP_PROP_FORMAT=-1 hands out the Annex-B packets of the bundled clip)."""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class CUVIDDECODECAPS(C.Structure):          # nvcuvid/cuviddec.h, restated (the header is not in the image)
    _fields_ = [("eCodecType", C.c_int), ("eChromaFormat", C.c_int), ("nBitDepthMinus8", C.c_uint), ("reserved1", C.c_uint * 3),
                ("bIsSupported", C.c_ubyte), ("nNumNVDECs", C.c_ubyte), ("nOutputFormatMask", C.c_ushort),
                ("nMaxWidth", C.c_uint), ("nMaxHeight", C.c_uint), ("nMaxMBCount", C.c_uint), ("nMinWidth", C.c_ushort),
                ("nMinHeight", C.c_ushort), ("bIsHistogramSupported", C.c_ubyte), ("nCounterBitDepth", C.c_ubyte),
                ("nMaxHistogramBins", C.c_ushort), ("reserved3", C.c_uint * 10)]


def main():
    out = []
    p = lambda *a: out.append(" ".join(str(x) for x in a))
    p("NVIDIA_DRIVER_CAPABILITIES =", os.environ.get("NVIDIA_DRIVER_CAPABILITIES"))
    libs = sorted(set(glob.glob("/usr/lib/x86_64-linux-gnu/libnvcuvid*") + glob.glob("/usr/lib64/libnvcuvid*") +
                      glob.glob("/usr/local/nvidia/lib64/libnvcuvid*") + glob.glob("/usr/lib/x86_64-linux-gnu/libnvidia-encode*")))
    p("video libraries on disk:", libs or "none")
    try:
        r = subprocess.run("ldconfig -p | grep -i -E 'nvcuvid|nvidia-encode|libcuda\\.so'", shell=True, capture_output=True, text=True)
        p("ldconfig:", r.stdout.strip() or "(no match)")
    except Exception as e:
        p("ldconfig failed:", e)
    lib = None
    for name in ["libnvcuvid.so.1", "libnvcuvid.so"] + libs:
        try:
            lib = C.CDLL(name)
            p("dlopen", name, "OK")
            break
        except OSError as e:
            p("dlopen", name, "failed:", e)
    if lib is not None:
        import torch
        torch.zeros(1, device="cuda")           # a current primary context
        for codec, cname in ((4, "H264"), (8, "HEVC"), (11, "AV1")):
            caps = CUVIDDECODECAPS()
            caps.eCodecType, caps.eChromaFormat, caps.nBitDepthMinus8 = codec, 1, 0      # cudaVideoChromaFormat_420
            rc = lib.cuvidGetDecoderCaps(C.byref(caps))
            p(f"cuvidGetDecoderCaps({cname} 4:2:0 8-bit) rc={rc} supported={caps.bIsSupported} nvdecs={caps.nNumNVDECs} "
              f"max={caps.nMaxWidth}x{caps.nMaxHeight} outmask={caps.nOutputFormatMask:#x}")
    try:
        import cv2
        clip = os.path.join(ROOT, "tests", "golden", "bundled_veo3_360p.mp4")
        cap = cv2.VideoCapture(clip, cv2.CAP_FFMPEG, [cv2.CAP_PROP_FORMAT, -1])
        n = tot = 0
        first = None
        while True:
            ok, pkt = cap.read()
            if not ok:
                break
            first = first if first is not None else bytes(pkt.ravel()[:8]).hex()
            n += 1
            tot += pkt.size
        p(f"OpenCV raw-packet demux of the bundled clip: {n} packets, {tot} bytes, first bytes {first} (Annex-B start code = 00000001)")
    except Exception as e:
        p("OpenCV raw-packet demux failed:", e)
    try:
        r = subprocess.run(["nvidia-smi", "--query-gpu=name,driver_version", "--format=csv,noheader"], capture_output=True, text=True)
        p("nvidia-smi:", r.stdout.strip())
    except Exception as e:
        p("nvidia-smi failed:", e)
    text = "\n".join(out)
    print(text)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "nvdec_probe.txt"), "w").write(text + "\n")


if __name__ == "__main__":
    main()
