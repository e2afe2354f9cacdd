/* Forced-include replacement for the reference's def.h (TEST INFRASTRUCTURE).
 * def.h:1-2 guards itself with _def_h_; pre-defining the guard lets the harness build the
 * unmodified reference sources for a different number of grid cells (def.h:10 hard-codes
 * 67420).  All other constants keep the reference's values (def.h:14-26). */
#ifndef _def_h_
#define _def_h_
#ifndef WGK_REF_NG
#error "WGK_REF_NG must be defined"
#endif
#define ng WGK_REF_NG
#define ng_climate 70412
#define nlct 18
#define reservoir_dsc 5
#define wateruse_dsc 5
#define ng_glolakcells 0
#endif
