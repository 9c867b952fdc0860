/* The three screen helpers of InOutFun/print_ScreenMessage.c, whose TU cannot be compiled without
 * PETSc headers (it includes <petscsys.h>).  Only needed when the reference is built without PETSc. */
#include <stdio.h>
int ResultsTimeStep;
void print_Status(char *Message, int Time) { (void)Time; puts(Message); }
void print_step(int Time, int NumTimeStep, double DeltaTimeStep) {
  printf("Step: [%i/%i] | DeltaT: %1.2e \n", Time, NumTimeStep, DeltaTimeStep);
}
void print_convergence_stats(int Time, int NumTimeStep, int Iter, int MaxIter, double Error0, double Error_total,
                             double Error_relative) {
  printf("Step [%i/%i] iter %i/%i err0 %e err %e rel %e\n", Time, NumTimeStep, Iter, MaxIter, Error0, Error_total,
         Error_relative);
}
