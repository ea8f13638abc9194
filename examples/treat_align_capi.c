/*
 * treat_align_capi.c -- `treat align -t templates.fa -f reads.fa` written against include/treat_b200.h alone, in plain C:
 * exactly the calls a cgo binding makes.  Template construction stays on the host
 * (NewTemplateFromFasta + SetOffset), the reads file goes to the GPU as bytes and the text comes back through a sink.
 *
 *   treat_align_capi -t templates.fa -f reads.fa [-b T] [--offset N] [--device D]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "treat_b200.h"

static int to_stdout(void *user, const uint8_t *text, uint64_t len, uint64_t first_record, uint64_t n_records)
{
    (void)first_record;
    (void)n_records;
    return fwrite(text, 1, (size_t)len, (FILE *)user) != (size_t)len;  /* non-zero stops the run */
}

int main(int argc, char **argv)
{
    const char *tpath = NULL, *fpath = NULL, *base = "T";
    int offset = 0, device = 0, i;
    for (i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "-t")) tpath = argv[i + 1];
        else if (!strcmp(argv[i], "-f")) fpath = argv[i + 1];
        else if (!strcmp(argv[i], "-b")) base = argv[i + 1];
        else if (!strcmp(argv[i], "--offset")) offset = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
    }
    if (!tpath || !fpath || strlen(base) != 1) {
        fprintf(stderr, "usage: treat_align_capi -t templates.fa -f reads.fa [-b T] [--offset N] [--device D]\n");
        return 2;
    }
    char err[512] = {0};
    treat_b200_template *tmpl = NULL;
    int rc = treat_b200_template_from_fasta(tpath, 1 /* FORWARD */, (uint8_t)base[0], offset, &tmpl, err, sizeof err);
    if (rc != TREAT_B200_OK) {
        fprintf(stderr, "FATA %s\n", err);
        return 1;
    }
    treat_b200_ctx *ctx = NULL;
    if ((rc = treat_b200_create(device, &ctx)) != TREAT_B200_OK) {
        fprintf(stderr, "FATA %s\n", treat_b200_strerror(rc));
        return 1;
    }
    if ((rc = treat_b200_use_template(ctx, tmpl)) != TREAT_B200_OK) {
        fprintf(stderr, "FATA %s\n", treat_b200_last_error(ctx));
        return 1;
    }
    FILE *fp = fopen(fpath, "rb");
    if (!fp) {
        fprintf(stderr, "FATA open %s: no such file or directory\n", fpath);
        return 1;
    }
    fseek(fp, 0, SEEK_END);
    long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    void *bytes = NULL;  /* pinned: the copies to the GPU run by DMA */
    if (size < 0 || treat_b200_host_alloc(&bytes, (size_t)size + 1) != TREAT_B200_OK) {
        fprintf(stderr, "FATA cannot read %s\n", fpath);
        return 1;
    }
    size_t got = fread(bytes, 1, (size_t)size, fp);
    fclose(fp);
    uint64_t n_records = 0;
    rc = treat_b200_fasta_text(ctx, (const uint8_t *)bytes, got, TREAT_B200_NO_SUMMARY, 80, to_stdout, stdout, &n_records);
    if (rc != TREAT_B200_OK) fprintf(stderr, "FATA %s: %s\n", treat_b200_strerror(rc), treat_b200_last_error(ctx));
    treat_b200_host_free(bytes);
    treat_b200_template_free(tmpl);
    treat_b200_destroy(ctx);
    return rc == TREAT_B200_OK ? 0 : 1;
}
