/*
 * ref_ensemble_params.c -- TEST INFRASTRUCTURE ONLY.
 *
 * resolve_run_params (lib/src/ensemble.c:55-76) and its table (:33-46) are static in the reference.  To pin the
 * product's kb200_ensemble_run_params against the reference's OWN code -- not against a second restatement --
 * this translation unit includes the reference's ensemble.c where it lies (nothing is copied into the
 * repository) and exports one wrapper.  Built by oracle/Makefile into oracle/_ref/libref_ensemble.so.
 */
#include "ensemble.c"

int refh_resolve_run_params(float base_gpo, float base_gpe, float base_tgpe, int k, uint64_t seed,
                            float* gpo, float* gpe, float* tgpe, uint64_t* run_seed, float* noise)
{
        resolve_run_params(base_gpo, base_gpe, base_tgpe, k, seed, gpo, gpe, tgpe, run_seed, noise);
        return 0;
}
