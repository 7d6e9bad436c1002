/*
 * deconvolute.c -- the C ABI of include/mdb200.h from plain C: what a compiled host (the
 * reference's Rust crate through metabodecon-sys, or any C/C++ program) does.
 *
 *   gcc -O2 -Iinclude examples/deconvolute.c -o examples/deconvolute \
 *       -Lmetabodecon_rust_b200 -lmdb200 -Wl,-rpath,$PWD/metabodecon_rust_b200 -lm
 *   ./examples/deconvolute tests/golden/bruker/blood_01/10/pdata/10/1r
 *
 * Reads a Bruker `1r` file (little-endian int32, NC_proc = 0, as blood_01), builds the axis the way
 * formats/bruker.rs:278-280 does, and runs Deconvoluter::default() with the water region ignored
 * (BASELINE config 1).  Prints the peak / Lorentzian counts and the MSE.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "mdb200.h"

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s path/to/1r [maximum_ppm width_ppm]\n", argv[0]);
        return 2;
    }
    const double maximum = argc > 2 ? atof(argv[2]) : 14.81146;
    const double width = argc > 3 ? atof(argv[3]) : 20.0236139622347;
    FILE *fh = fopen(argv[1], "rb");
    if (!fh) { perror(argv[1]); return 2; }
    fseek(fh, 0, SEEK_END);
    const size_t n = (size_t)ftell(fh) / 4;
    fseek(fh, 0, SEEK_SET);
    int32_t *raw = malloc(n * 4);
    double *x = malloc(n * 8), *y = malloc(n * 8);
    if (fread(raw, 4, n, fh) != n) { fprintf(stderr, "short read\n"); return 2; }
    fclose(fh);
    for (size_t i = 0; i < n; ++i) {
        x[i] = maximum - (double)i * width / ((double)n - 1.0);
        y[i] = (double)raw[i];
    }

    if (mdb_device_count() < 1) {
        fprintf(stderr, "no CUDA device: libmdb200 has no CPU fallback\n");
        return 3;
    }
    const double sb[2] = {-2.2, 11.8};
    double ordered[2];
    mdb_status st = mdb_spectrum_validate(x, n, y, n, sb, ordered); /* Spectrum::new */
    if (st != MDB_OK) { fprintf(stderr, "invalid spectrum: %s\n", mdb_last_error_message()); return 1; }

    /* Peak sets and Lorentzians are the reference's bit patterns in either mode; the MSE and
     * superposition values are too under MDB_SUPERPOSITION=exact (or
     * mdb_set_superposition_mode(MDB_SUPERPOSITION_EXACT)), and within about 1e-15 (MSE: 1e-13) relative of them
     * in the default mode. */
    mdb_deconvoluter *dec = NULL;
    mdb_deconvoluter_default(&dec);
    mdb_deconvoluter_add_ignore_region(dec, 4.7, 4.9);

    mdb_spectrum_view view = {x, y, n, {ordered[0], ordered[1]}};
    mdb_batch *batch = NULL;
    st = mdb_deconvolute_spectra(dec, &view, 1, MDB_MEM_HOST, &batch);
    if (st != MDB_OK) {
        fprintf(stderr, "deconvolution failed (%d): %s\n", (int)st, mdb_last_error_message());
        return 1;
    }
    const size_t k = mdb_batch_n_lorentzians(batch, 0);
    const mdb_lorentzian *lor = mdb_batch_lorentzians(batch, 0);
    printf("points %zu selected_peaks %zu lorentzians %zu mse %a\n", n, mdb_batch_n_peaks(batch, 0), k,
           mdb_batch_mse(batch, 0));
    for (size_t i = 0; i < 3 && i < k; ++i)
        printf("L[%zu] sfhw %a hw2 %a maxp %a\n", i, lor[i].sfhw, lor[i].hw2, lor[i].maxp);

    /* Lorentzian::superposition_vec on the first 8 axis points */
    double sup[8];
    st = mdb_superposition_vec(x, 8, lor, k, sup, MDB_MEM_HOST);
    if (st != MDB_OK) { fprintf(stderr, "superposition failed: %s\n", mdb_last_error_message()); return 1; }
    printf("superposition(x[0]) %a\n", sup[0]);

    mdb_batch_free(batch);
    mdb_deconvoluter_free(dec);
    free(raw); free(x); free(y);
    return 0;
}
