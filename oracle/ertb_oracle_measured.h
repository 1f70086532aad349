/* ertb_oracle_measured.h -- measured_mono BSDF restatement (oracle side). See ertb_oracle_measured.c */
#ifndef ERTB_ORACLE_MEASURED_H
#define ERTB_ORACLE_MEASURED_H
/* `table` = ertb_scene_desc::bsdf_table of a measured_mono scene; local-frame vectors; BSDF::eval / pdf / sample
 * exactly as the plugin returns them (eval carries the cosine; sample returns eval / pdf) */
double mm_oracle_eval(const float *table, const double wi[3], const double wo[3]);
double mm_oracle_pdf(const float *table, const double wi[3], const double wo[3]);
double mm_oracle_sample(const float *table, const double wi[3], double u1, double u2, double wo[3], double *pdf);
#endif
