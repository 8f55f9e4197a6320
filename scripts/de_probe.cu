// probe: does this B200 + driver expose the hardware decompression engine through the driver API?
#include <cuda.h>
#include <stdio.h>
int main()
{
    cuInit(0);
    CUdevice d;
    cuDeviceGet(&d, 0);
    int mask = -1, maxlen = -1;
    CUresult r1 = cuDeviceGetAttribute(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, d);
    CUresult r2 = cuDeviceGetAttribute(&maxlen, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, d);
    printf("decompress algorithm mask rc=%d mask=0x%x (1=deflate 2=snappy 4=lz4)  max length rc=%d %d\n", (int)r1, mask, (int)r2, maxlen);
    return 0;
}
