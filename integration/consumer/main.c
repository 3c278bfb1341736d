/* consumer of kalign::kalign: aligns a few sequences through the public kalign() of
 * lib/include/kalign/kalign.h:45 and prints the rows.  Exit code 0 = aligned, 3 = kalign() failed
 * (what happens without a CUDA device: the GPU build has no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <kalign/kalign.h>

int main(int argc, char** argv)
{
        char* in[4] = {"GKGDPKKPRGKMSSYAFFVQTSREEHKKKHPDASVNFSEFSKKCSERWKTMSAKEKGKFEDMAKADKARYEREMKTYIPPKGE",
                       "MQDRVKRPMNAFIVWSRDQRRKMALENPRMRNSEISKQLGYQWKMLTEAEKWPFFQEAQKLQAMHREKYPNYKYRPRRKAKMLPK",
                       "MKKLKKHPDFPKKPLTPYFRFFMEKRAKYAKLHPEMSNLDLTKILSKKYKELPEKKKMKYIQDFQREKQEFERNLARFREDHPDLIQNAKK",
                       "MHIKKPLNAFMLYMKEMRANVVAESTLKESAAINQILGRRWHALSREEQAKYYELARKERQLHMQLYPGWSARDNYGKKKKRKREK"};
        int len[4];
        char** aligned = NULL;
        int alnlen = 0;
        int i;
        (void)argc; (void)argv;
        for(i = 0; i < 4; i++){
                len[i] = (int)strlen(in[i]);
        }
        if(kalign(in, len, 4, 2, KALIGN_TYPE_PROTEIN, -1.0f, -1.0f, -1.0f, &aligned, &alnlen) != 0){
                fprintf(stderr, "kalign() failed\n");
                return 3;
        }
        for(i = 0; i < 4; i++){
                printf("%s\n", aligned[i]);
                free(aligned[i]);
        }
        free(aligned);
        return 0;
}
