// Probe: which codecs this GPU's NVDEC decodes (cuvidGetDecoderCaps through a dlopen'ed libnvcuvid; the Video Codec SDK
// headers are not in the image, so the few prototypes used are declared by hand).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

typedef struct {
    int eCodecType; int eChromaFormat; unsigned int nBitDepthMinus8; unsigned int reserved1[3];
    unsigned char bIsSupported; unsigned char nNumNVDECs; unsigned short nOutputFormatMask;
    unsigned int nMaxWidth; unsigned int nMaxHeight; unsigned int nMaxMBCount; unsigned short nMinWidth; unsigned short nMinHeight;
    unsigned char bIsHistogramSupported; unsigned char nCounterBitDepth; unsigned short nMaxHistogramBins; unsigned int reserved3[10];
} CUVIDDECODECAPS;
typedef int (*caps_fn)(CUVIDDECODECAPS*);

int main() {
    cudaFree(0);
    void* h = dlopen("libnvcuvid.so.1", RTLD_NOW);
    if (!h) { printf("no libnvcuvid: %s\n", dlerror()); return 0; }
    caps_fn caps = (caps_fn)dlsym(h, "cuvidGetDecoderCaps");
    if (!caps) { printf("no cuvidGetDecoderCaps\n"); return 0; }
    const char* names[] = {"MPEG1", "MPEG2", "MPEG4", "VC1", "H264", "JPEG", "H264_SVC", "H264_MVC", "HEVC", "VP8", "VP9", "AV1"};
    printf("sizeof caps %zu\n", sizeof(CUVIDDECODECAPS));
    for (int c = 0; c < 12; ++c) {
        CUVIDDECODECAPS k; memset(&k, 0, sizeof k);
        k.eCodecType = c; k.eChromaFormat = 1; k.nBitDepthMinus8 = 0;
        int rc = caps(&k);
        printf("%-8s rc=%d supported=%d nvdecs=%d outmask=0x%x max %ux%u min %ux%u mb %u\n", names[c], rc, k.bIsSupported, k.nNumNVDECs,
               k.nOutputFormatMask, k.nMaxWidth, k.nMaxHeight, k.nMinWidth, k.nMinHeight, k.nMaxMBCount);
    }
    return 0;
}
