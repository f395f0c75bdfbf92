// TEST INFRASTRUCTURE ONLY (oracle/).  The reference programs declare `void main`, which g++
// rejects, so they are compiled with -Dmain=ref_main and entered through this 2-line driver.
void ref_main(int, char **);
int main(int argc, char **argv) { ref_main(argc, argv); return 0; }
