/* Out-of-process driver for the reference's compiled ctypes builders (oracle/_ref/<name>/run.so).
 * TEST INFRASTRUCTURE ONLY.  Loading those libraries into a Python process that already imported numpy
 * crashes inside libstdc++ (symbol interposition), so they are run here, isolated:
 *   ref_driver tet   <run.so> <in.bin> <out.bin>   in: int32 n_point, n_tet, out_cap_ints, tet[n_tet*4]
 *                                                   out: int32 n_out, then out_cap_ints int32
 *   ref_driver colaps <run.so> <in.bin> <out.bin>  in: int32 n_point, float pts[n_point*3]
 *                                                   out: int32 n_unique, map[n_point], inverse[n_point]   */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void* slurp(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    *n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    void* p = malloc(*n ? *n : 1);
    if (fread(p, 1, *n, f) != *n) { fclose(f); return NULL; }
    fclose(f);
    return p;
}

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: ref_driver tet|colaps run.so in.bin out.bin\n"); return 2; }
    void* h = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "%s\n", dlerror()); return 3; }
    void* sym = dlsym(h, "run");
    if (!sym) { fprintf(stderr, "no symbol run\n"); return 3; }
    size_t n = 0;
    char* in = (char*)slurp(argv[3], &n);
    if (!in) { fprintf(stderr, "cannot read %s\n", argv[3]); return 4; }
    FILE* out = fopen(argv[4], "wb");
    if (!out) return 4;
    if (strcmp(argv[1], "tet") == 0) {
        int32_t* hdr = (int32_t*)in;
        int n_point = hdr[0], n_tet = hdr[1];
        size_t cap = (size_t)(uint32_t)hdr[2];
        int32_t* buf = (int32_t*)calloc(cap ? cap : 1, 4);
        int32_t n_out = 0;
        ((void (*)(int*, int*, int*, int, int))sym)(hdr + 3, buf, &n_out, n_point, n_tet);
        fwrite(&n_out, 4, 1, out);
        fwrite(buf, 4, cap, out);
    } else {
        int32_t n_point = *(int32_t*)in;
        int32_t* map = (int32_t*)calloc(n_point ? n_point : 1, 4);
        int32_t* inv = (int32_t*)calloc(n_point ? n_point : 1, 4);
        int32_t cnt = 0;
        ((void (*)(float*, int*, int*, int*, int))sym)((float*)(in + 4), map, inv, &cnt, n_point);
        fwrite(&cnt, 4, 1, out);
        fwrite(map, 4, n_point, out);
        fwrite(inv, 4, n_point, out);
    }
    fclose(out);
    return 0;
}
