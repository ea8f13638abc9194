/*
 * oracle/oracle_cli.c -- `treat align` restated on top of the CPU oracle
 * (cmd/treat/main.go:102-111 flags, cmd/treat/align.go:59-122 behaviour).
 * TEST INFRASTRUCTURE ONLY.
 *   oracle_align [-t templates.fa] [-f reads.fa] [-1 S1 -2 S2] [-b T] [--offset N] [--tie ULD]
 */
#include "treat_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    const char *t = NULL, *f = NULL, *s1 = NULL, *s2 = NULL, *base = "T";
    int offset = 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        const char *v = (i + 1 < argc) ? argv[i + 1] : NULL;
        if ((!strcmp(a, "-t") || !strcmp(a, "--template")) && v) t = v, i++;
        else if ((!strcmp(a, "-f") || !strcmp(a, "--fragment")) && v) f = v, i++;
        else if ((!strcmp(a, "-1") || !strcmp(a, "--s1")) && v) s1 = v, i++;
        else if ((!strcmp(a, "-2") || !strcmp(a, "--s2")) && v) s2 = v, i++;
        else if ((!strcmp(a, "-b") || !strcmp(a, "--base")) && v) base = v, i++;
        else if (!strcmp(a, "--offset") && v) offset = atoi(v), i++;
        else if (!strcmp(a, "--tie") && v) {
            const char *names[] = {"ULD", "UDL", "LUD", "LDU", "DUL", "DLU"};
            for (int k = 0; k < 6; k++)
                if (!strcmp(v, names[k])) orc_set_tie_order(k);
            i++;
        } else {
            fprintf(stderr, "unknown argument %s\n", a);
            return 2;
        }
    }
    if (strlen(base) != 1) {
        fprintf(stderr, "Please provide the edit base\n");
        return 1;
    }
    char err[512] = {0};
    size_t n = 0;
    char *out = orc_cli_align(t, f, s1, s2, base[0], offset, &n, err, sizeof err);
    if (!out) {
        fprintf(stderr, "%s\n", err);
        return 1;
    }
    fwrite(out, 1, n, stdout);
    orc_free(out);
    return 0;
}
