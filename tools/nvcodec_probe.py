#!/usr/bin/env python
"""SURVEY.md 8f-3 probe: are NVDEC / NVENC usable on the GPU box?  dlopen the driver's codec libraries and ask
NVDEC for its H.264 / HEVC / AV1 / VP9 decode caps; list the PyAV / ffmpeg / cv2 pieces a decode path would need."""
import ctypes, json, shutil, sys
out = {}
for name in ("libnvcuvid.so.1", "libnvcuvid.so", "libnvidia-encode.so.1", "libnvidia-encode.so"):
    try:
        ctypes.CDLL(name)
        out[name] = "loaded"
    except OSError as e:
        out[name] = f"absent ({str(e)[:80]})"
out["ffmpeg_binary"] = shutil.which("ffmpeg")
out["ffprobe_binary"] = shutil.which("ffprobe")
for mod in ("av", "PyNvVideoCodec", "nvidia.dali", "torchvision.io", "torchcodec", "cv2.cudacodec"):
    try:
        __import__(mod)
        out["import " + mod] = "ok"
    except Exception as e:
        out["import " + mod] = f"no ({type(e).__name__})"
try:
    import cv2
    info = cv2.getBuildInformation()
    out["cv2_ffmpeg"] = [l.strip() for l in info.splitlines() if "FFMPEG" in l or "NVCUVID" in l or "CUDA" in l][:6]
except Exception as e:
    out["cv2"] = str(e)
# NVDEC capabilities through cuvidGetDecoderCaps
try:
    import torch
    torch.cuda.init(); torch.zeros(1).cuda()
    lib = ctypes.CDLL("libnvcuvid.so.1")
    class CAPS(ctypes.Structure):
        _fields_ = [("eCodecType", ctypes.c_int), ("eChromaFormat", ctypes.c_int), ("nBitDepthMinus8", ctypes.c_uint),
                    ("reserved1", ctypes.c_uint * 3), ("bIsSupported", ctypes.c_ubyte), ("nNumNVDECs", ctypes.c_ubyte),
                    ("nOutputFormatMask", ctypes.c_ushort), ("nMaxWidth", ctypes.c_uint), ("nMaxHeight", ctypes.c_uint),
                    ("nMaxMBCount", ctypes.c_uint), ("nMinWidth", ctypes.c_ushort), ("nMinHeight", ctypes.c_ushort),
                    ("bIsHistogramSupported", ctypes.c_ubyte), ("nCounterBitDepth", ctypes.c_ubyte),
                    ("nMaxHistogramBins", ctypes.c_ushort), ("reserved3", ctypes.c_uint * 10)]
    for codec, cid in (("h264", 4), ("hevc", 8), ("vp9", 10), ("av1", 11)):
        c = CAPS(); c.eCodecType = cid; c.eChromaFormat = 1; c.nBitDepthMinus8 = 0
        r = lib.cuvidGetDecoderCaps(ctypes.byref(c))
        out["nvdec_" + codec] = {"rc": r, "supported": int(c.bIsSupported), "engines": int(c.nNumNVDECs),
                                 "max": [int(c.nMaxWidth), int(c.nMaxHeight)]}
except Exception as e:
    out["nvdec_caps"] = f"{type(e).__name__}: {e}"
print(json.dumps(out, indent=1))
