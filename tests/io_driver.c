/* io_driver.c -- TEST INFRASTRUCTURE: dumps what kalign_read_input (lib/include/kalign/kalign.h:36) made of a file
   and, when the input is an alignment, writes it back with kalign_write_msa.  Linked once against the drop-in
   library (integration/_out/libkalign.so.3) and once against the unmodified reference (oracle/_ref); the test
   compares the two outputs byte for byte.  usage: io_driver in.fa [out.afa [format]] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "msa_struct.h"

int kalign_read_input(char* infile, struct msa** msa, int quiet);
int kalign_write_msa(struct msa* msa, char* outfile, char* format);
int finalise_alignment(struct msa* msa);
void kalign_free_msa(struct msa* msa);

int main(int argc, char** argv)
{
        struct msa* msa = NULL;
        int rc;
        if(argc < 2){
                return 2;
        }
        rc = kalign_read_input(argv[1], &msa, 1);
        printf("rc %d msa %s\n", rc, msa ? "yes" : "null");
        if(rc != 0 || !msa){
                return 0;
        }
        printf("numseq %d alloc %d biotype %d aligned %d L %d num_profiles %d\n", msa->numseq, msa->alloc_numseq, msa->biotype,
               msa->aligned, msa->L, msa->num_profiles);
        for(int i = 0; i < 128; i++){
                if(msa->letter_freq[i]){
                        printf("freq %d %d\n", i, msa->letter_freq[i]);
                }
        }
        for(int i = 0; i < msa->numseq; i++){
                struct msa_seq* s = msa->sequences[i];
                printf("seq %d name [%s] len %d alloc %d rank %d\n%s\ngaps", i, s->name, s->len, s->alloc_len, s->rank, s->seq);
                for(int j = 0; j <= s->len; j++){
                        if(s->gaps[j]){
                                printf(" %d:%d", j, s->gaps[j]);
                        }
                }
                printf("\n");
        }
        if(argc >= 3 && msa->aligned == ALN_STATUS_ALIGNED){
                rc = finalise_alignment(msa);
                printf("finalise %d alnlen %d\n", rc, msa->alnlen);
                if(rc == 0){
                        rc = kalign_write_msa(msa, argv[2], argc >= 4 ? argv[3] : NULL);
                        printf("write %d\n", rc);
                }
        }
        kalign_free_msa(msa);
        return 0;
}
