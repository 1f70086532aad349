/* placeholder, replaced below */
#include "ertb_oracle_ocean.h"
#include <string.h>
int ocean_init(ocean_state_t *o, const float *p) { (void)p; memset(o, 0, sizeof *o); return 1; }
void ocean_free(ocean_state_t *o) { (void)o; }
double ocean_eval(const ocean_state_t *o, double a, double b, double c, double d, double e, double f) { (void)o;(void)a;(void)b;(void)c;(void)d;(void)e;(void)f; return 0; }
double ocean_sample(const ocean_state_t *o, double a, double b, double c, double s, double u, double v, double *wo) { (void)o;(void)a;(void)b;(void)c;(void)s;(void)u;(void)v; wo[0]=wo[1]=0; wo[2]=1; return 0; }
double ocean_pdf(const ocean_state_t *o, double a, double b, double c, double d, double e, double f) { (void)o;(void)a;(void)b;(void)c;(void)d;(void)e;(void)f; return 0; }
